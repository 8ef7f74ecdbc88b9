#!/usr/bin/env python
"""Random-action rollout on the b200sim backend -- the counterpart of mani_skill/examples/demo_random_action.py (same short flags).

    python examples/demo_random_action.py -e PickCube-v1 -n 1024 -o state
    python examples/demo_random_action.py -e PegInsertionSide-v1 -n 4 -o rgbd --record-dir out/     # trajectory .npz/.json + .mp4 per episode batch

Runs until every sub-scene has finished one episode (success or the task's time limit) and prints the return / success statistics the
vector wrapper keeps (mani_skill/vector/wrappers/gymnasium.py:131-156)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maniskill_b200 as ms  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-e", "--env-id", default="PickCube-v1", choices=sorted(ms.REGISTERED_ENVS))
    ap.add_argument("-o", "--obs-mode", default="state")
    ap.add_argument("-n", "--num-envs", type=int, default=1)
    ap.add_argument("-c", "--control-mode", default=None)
    ap.add_argument("--reward-mode", default=None)
    ap.add_argument("--render-mode", default="rgb_array", help="rgb_array | sensors | all | none")
    ap.add_argument("--record-dir", default=None, help="save trajectories (and videos, when a render mode is set) here")
    ap.add_argument("-s", "--seed", type=int, default=None)
    ap.add_argument("--quiet", action="store_true")
    args = ap.parse_args()
    render_mode = None if args.render_mode == "none" else args.render_mode
    env = ms.make(args.env_id, num_envs=args.num_envs, obs_mode=args.obs_mode, control_mode=args.control_mode, reward_mode=args.reward_mode,
                  render_mode=render_mode)
    if args.record_dir:
        env = ms.RecordEpisode(env, args.record_dir, save_video=render_mode is not None, max_steps_per_video=env.max_episode_steps)
    venv = ms.ManiSkillVectorEnv(env, auto_reset=True, record_metrics=True)
    if not args.quiet:
        base = venv.base_env
        print(f"{args.env_id}: {args.num_envs} sub-scenes on {venv.device}, control mode {base.control_mode}, action dim {base.action_dim}, "
              f"max_episode_steps {venv.max_episode_steps}")
    if args.seed is not None:
        torch.manual_seed(args.seed)
    obs, _ = venv.reset(seed=args.seed)
    finished = torch.zeros(args.num_envs, dtype=torch.bool, device=venv.device)
    returns, success, steps = torch.zeros(args.num_envs, device=venv.device), torch.zeros(args.num_envs, dtype=torch.bool, device=venv.device), 0
    while not bool(finished.all()):
        action = 2 * torch.rand((args.num_envs, venv.base_env.action_dim), device=venv.device) - 1
        obs, rew, terminated, truncated, info = venv.step(action)
        steps += 1
        done = terminated | truncated
        if "final_info" in info:
            ep = info["final_info"]["episode"]
            new = done & ~finished
            returns[new], success[new] = ep["return"][new], ep["success_once"][new]
            finished |= done
    if not args.quiet:
        print(f"{steps} control steps; mean return {returns.mean().item():.3f}, success rate {success.float().mean().item():.3f}")
    if args.record_dir:
        env.close()
        print("recorded to", args.record_dir)
    else:
        venv.close()


if __name__ == "__main__":
    main()
