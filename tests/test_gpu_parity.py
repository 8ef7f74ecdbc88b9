"""-m gpu: the CUDA path (through the C-ABI / ctypes) against the CPU oracle on the same seeded inputs.

north_star tolerance: joint q / qdot and actor poses within 1e-4 relative after 100 substeps.
"""
import numpy as np
import pytest

from scenarios import rel_err_vec, run_pick_cube

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4  # BASELINE.json north_star: "within 1e-4 rel after 100 substeps"


@pytest.fixture(scope="module")
def cm():
    from maniskill_b200.scenes import pick_cube_scene
    return pick_cube_scene(64).compile()


def test_pick_cube_100_substeps_vs_oracle_f32(cm):
    ref = run_pick_cube("oracle32", cm, 100)
    got = run_pick_cube("cuda", cm, 100)
    assert np.isfinite(got["body"]).all()
    e_q = rel_err_vec(got["qpos"], ref["qpos"], 0.1)
    e_qd = rel_err_vec(got["qvel"], ref["qvel"], 0.1)
    e_pos = rel_err_vec(got["body"][..., :3], ref["body"][..., :3], 0.1)
    e_quat = rel_err_vec(got["body"][..., 3:7], ref["body"][..., 3:7], 1.0)
    print("rel err q %.3g qd %.3g pos %.3g quat %.3g" % (e_q, e_qd, e_pos, e_quat))
    assert e_q < REL_TOL and e_qd < REL_TOL and e_pos < REL_TOL and e_quat < REL_TOL


def test_pick_cube_vs_oracle_f64_is_close(cm):
    """float32 kernel vs float64 oracle: looser, documents the precision of the fp32 path."""
    ref = run_pick_cube("oracle64", cm, 100)
    got = run_pick_cube("cuda", cm, 100)
    # chaotic contact events can flip in a few envs; compare the median env
    err = np.abs(got["qpos"] - ref["qpos"]).max(axis=1)
    print("median |dq| vs f64", np.median(err), "max", err.max())
    assert np.median(err) < 1e-3


def test_contact_query_matches_oracle(cm):
    ref = run_pick_cube("oracle32", cm, 60)
    got = run_pick_cube("cuda", cm, 60)
    w = got["world"]
    rows = cm.link_rows["panda"]
    cube = cm.actor_rows["cube"]
    table = cm.actor_rows["table-workspace"]
    from maniskill_b200.backend import ANY_BODY
    key = w.create_contact_query([(rows["panda_leftfinger"], cube), (rows["panda_rightfinger"], cube), (cube, table), (cube, ANY_BODY),
                                  (rows["panda_link7"], ANY_BODY)])
    import torch
    imp = w.query_contact_impulses(key)
    torch.cuda.synchronize()
    imp = imp.cpu().numpy()
    o = ref["world"]
    ref_imp = np.stack([o.pair_impulse(rows["panda_leftfinger"], cube), o.pair_impulse(rows["panda_rightfinger"], cube),
                        o.pair_impulse(cube, table), o.pair_impulse(cube, ANY_BODY), o.pair_impulse(rows["panda_link7"], ANY_BODY)], axis=1)
    # the body-net query (px.gpu_query_contact_body_impulses) is the sum of the pairwise ones over everything the body touches
    assert np.abs(ref_imp[:, 3] - (ref_imp[:, 2] - ref_imp[:, 0] - ref_imp[:, 1])).max() < 1e-5
    assert np.abs(imp - ref_imp).max() < 1e-4 * max(1.0, np.abs(ref_imp).max())
    # the cube rests on the table in most envs: vertical impulse = m g dt
    assert np.median(np.abs(ref_imp[:, 2, 2])) == pytest.approx(0.064 * 9.81 * 0.01, rel=0.05)


@pytest.mark.parametrize("idx", [0, 1, 2, 3, 4], ids=["bodies-only", "articulation-only", "two-bodies-no-static", "hull-on-box", "hull-on-hull"])
def test_edge_models_vs_oracle(idx):
    """Degenerate model shapes through the C-ABI: 0 dofs, 0 candidate pairs / 0 bodies, bodies without static geometry
    (the scenes of tests/test_emu_parity.py::test_emu_edge_models_match_oracle)."""
    import torch

    from maniskill_b200.backend import BUF_ALL, World
    from oracle.oracle import OracleWorld
    from test_emu_parity import _edge_scenes
    name, scene, n_sub = _edge_scenes()[idx]
    cm = scene.compile()
    o, w = OracleWorld(cm, "f32"), World(cm)
    nl = cm.scalars["n_link"]
    if idx == 1:
        tq = np.tile(np.array([[0.4, 0.02]]), (cm.scalars["n_envs"], 1))
        o.set_joint("target_qpos", tq)
        w.target_qpos[:] = torch.tensor(tq, dtype=torch.float32, device=w.device)
        w.apply(32)
    if idx == 2:
        b = o.get_bodies()
        b[:, 0, 7], b[:, 1, 7] = 1.0, -0.5
        o.set_bodies(b)
        w.body_view()[:, nl:] = torch.tensor(b, dtype=torch.float32, device=w.device)
        w.apply()
    o.step(n_sub)
    for _ in range(n_sub // 5):
        w.step(5, 0)
    w.fetch(BUF_ALL)
    torch.cuda.synchronize()
    ref, got = o.rigid_body_data(), w.body_view().double().cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got[..., :7] - ref[..., :7]).max() < 1e-4, np.abs(got[..., :7] - ref[..., :7]).max()
    if cm.scalars["n_dof"]:
        assert np.abs(w.qpos.double().cpu().numpy() - o.get_joint("qpos")).max() < 1e-4
    assert int(w.overflow_flag.item()) == 0
    w.close()
