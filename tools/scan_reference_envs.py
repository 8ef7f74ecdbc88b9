"""Which of the reference's registered environments build, reset and step UNCHANGED on the shim?  (DESIGN.md section 4.3 quotes the result: 36 of 74.)

    python tools/scan_reference_envs.py [--reference /root/reference] [--emu]

`--emu`: the host emulation of the device code (tests/emu) instead of the CUDA world -- what a machine without a GPU can run.  Every id is made with num_envs=2,
obs_mode="state", reset(seed=0) and stepped once; an asset-download prompt counts as "assets".  Stand-in PartNet cabinets are generated into a temporary
MS_ASSET_DIR unless the variable is set."""
import builtins
import os
import signal
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "tools")]


def main(argv):
    ref = argv[argv.index("--reference") + 1] if "--reference" in argv else ("/root/reference" if os.path.isdir("/root/reference/mani_skill") else os.path.join(ROOT, "baseline", "_ref"))
    emu = "--emu" in argv
    import torch

    import maniskill_b200.compat as compat
    compat.install()
    if "MS_ASSET_DIR" not in os.environ:
        import make_standin_partnet
        assets = tempfile.mkdtemp(prefix="b200sim_ms_assets_")
        make_standin_partnet.main(os.path.join(ref, "mani_skill", "assets", "partnet_mobility", "meta"), assets)
        os.environ["MS_ASSET_DIR"] = assets
    sys.path.insert(0, ref)
    import gymnasium as gym
    import mani_skill.envs  # noqa: F401
    if emu:
        from emu_world import EmuBackendWorld
        compat.WORLD_FACTORY = lambda cm, dev: EmuBackendWorld(cm)
        import mani_skill.envs.sapien_env as SE
        import mani_skill.envs.utils.system.backend as B
        orig = B.parse_sim_and_render_backend

        def parse(sim_backend, render_backend):
            info = orig(sim_backend, render_backend)
            info.device = torch.device("cpu")
            return info
        SE.parse_sim_and_render_backend = parse
        torch.cuda.synchronize = lambda *a, **k: None
    dev = torch.device("cpu") if emu else torch.device("cuda")
    from mani_skill.utils.registration import REGISTERED_ENVS
    builtins.input = lambda *a, **k: "n"     # "download the assets now?"

    class Timeout(Exception):
        pass

    def on_alarm(signum, frame):
        raise Timeout()
    signal.signal(signal.SIGALRM, on_alarm)
    ok, bad = [], {}
    for eid in REGISTERED_ENVS:
        signal.alarm(120)
        try:
            env = gym.make(eid, num_envs=2, obs_mode="state", sim_backend="physx_cuda")
            env.reset(seed=0)
            a = env.action_space.sample()
            a = {k: torch.as_tensor(v, device=dev) for k, v in a.items()} if isinstance(a, dict) else torch.as_tensor(a, device=dev)
            env.step(a)
            env.close()
            ok.append(eid)
        except SystemExit:
            bad[eid] = "assets"
        except BaseException as e:  # noqa: BLE001
            bad[eid] = f"{type(e).__name__}: {str(e)[:120]}"
        finally:
            signal.alarm(0)
    print(f"OK {len(ok)} of {len(REGISTERED_ENVS)}: {ok}")
    for k, v in bad.items():
        print(f"NO {k}: {v}")


if __name__ == "__main__":
    main(sys.argv[1:])
