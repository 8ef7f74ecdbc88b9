"""Runs the reference's OWN benchmark script (mani_skill/examples/benchmarking/gpu_sim.py -- the protocol BASELINE.json's metric is defined by) UNMODIFIED on the
b200sim backend: installs the `sapien` shim, puts the reference package on sys.path and hands the command line to its `main`.

    python examples/run_reference_benchmark.py [--reference <checkout or baseline/_ref>] -e PickCube-v1 -n 4096 -o state --sim-freq 100 --control-freq 20

What it prints is the throughput of the reference's python layer (its own controllers, observation / reward code: ~100 torch launches per step) over this backend's
`px.step()` / `gpu_fetch_*` / camera group.  `bench.py` times the fused path of this repo (the same task behind `ManiSkillVectorEnv`: one controller kernel, the
captured physics graph, one epilogue kernel) -- the two numbers bracket what a user of the reference gets before and after switching the env class.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv):
    ref = None
    if "--reference" in argv:
        i = argv.index("--reference")
        ref = argv[i + 1]
        argv = argv[:i] + argv[i + 2:]
    for cand in ([ref] if ref else []) + [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]:
        if cand and os.path.isdir(os.path.join(cand, "mani_skill")):
            ref = cand
            break
    else:
        raise SystemExit("no reference package found: pass --reference <dir that holds mani_skill/> or run tools/install_reference.py")
    import maniskill_b200.compat as compat
    compat.install()
    sys.path.insert(0, ref)
    import tyro
    from mani_skill.examples.benchmarking.gpu_sim import Args, main as reference_main
    reference_main(tyro.cli(Args, args=argv))


if __name__ == "__main__":
    main(sys.argv[1:])
