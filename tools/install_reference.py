"""Installs the UNMODIFIED reference python package into baseline/_ref (git-ignored; travels to the GPU box with the snapshot) so that the `-m gpu` test
tests/test_zzz_gpu_reference_package.py and bench.py's `reference_python_on_shim` side key can run haosulab/ManiSkill's own `mani_skill` on a B200 through the
`sapien` shim (maniskill_b200/compat).  This is NOT bench.py's `--impl reference` arm (SAPIEN / PhysX cannot be installed: that arm times the CPU oracle).

    pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>

(`/root/reference` is read-only and the build writes an egg-info next to setup.py: installed from a copy under /tmp; `--no-deps` because sapien, gymnasium, ... are
not installable here -- the shim provides them.)  After the install the asset directories of robots / environment maps no test uses are deleted to keep the
snapshot small (about 70 MB instead of 228 MB); every python file is byte-identical to the reference's.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref")
KEEP_ROBOTS = {"panda", "fetch", "so100", "humanoid", "hopper", "cartpole", "ant", "widowx", "xarm6", "floating_panda_gripper", "googlerobot"}


def install(reference="/root/reference", force=False) -> str:
    marker = os.path.join(TARGET, "mani_skill", "__init__.py")
    if os.path.exists(marker) and not force:
        return TARGET
    if not os.path.isdir(os.path.join(reference, "mani_skill")):
        raise RuntimeError(f"no reference checkout at {reference}")
    shutil.rmtree(TARGET, ignore_errors=True)
    os.makedirs(TARGET, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="b200sim_ref_") as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(reference, src, ignore=shutil.ignore_patterns(".git", "docs", "figures"))
        subprocess.check_call([sys.executable, "-m", "pip", "install", "-q", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
                               "--target", TARGET, src], cwd=tmp)
    # small data files the wheel leaves out (its package_data globs miss e.g. examples/benchmarking/envs/maniskill/assets/cartpole.xml): copied as they are
    src_pkg, dst_pkg = os.path.join(reference, "mani_skill"), os.path.join(TARGET, "mani_skill")
    for r, _, fs in os.walk(src_pkg):
        for f in fs:
            a = os.path.join(r, f)
            b = os.path.join(dst_pkg, os.path.relpath(a, src_pkg))
            if not os.path.exists(b) and not f.endswith((".pyc",)) and os.path.getsize(a) < (1 << 20) and "__pycache__" not in a:
                os.makedirs(os.path.dirname(b), exist_ok=True)
                shutil.copyfile(a, b)
    assets = os.path.join(TARGET, "mani_skill", "assets")
    shutil.rmtree(os.path.join(assets, "environment_maps"), ignore_errors=True)
    robots = os.path.join(assets, "robots")
    for d in os.listdir(robots):
        p = os.path.join(robots, d)
        if os.path.isdir(p) and d not in KEEP_ROBOTS and sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(p) for f in fs) > 2 << 20:
            shutil.rmtree(p)
    # single large data files nothing in the tests reads (python files are never touched)
    keep = ("robots/panda/franka_description", "robots/panda/realsense2_description/meshes/d415.stl", "robots/fetch", "utils/scene_builder/table",
            "robots/so100", "partnet_mobility")
    pkg = os.path.join(TARGET, "mani_skill")
    for sub in ("assets", os.path.join("utils", "scene_builder")):
        for r, _, fs in os.walk(os.path.join(pkg, sub)):
            for f in fs:
                path = os.path.join(r, f)
                rel = os.path.relpath(path, pkg).replace(os.sep, "/")
                if not f.endswith(".py") and os.path.getsize(path) > (1 << 20) and not any(k in rel for k in keep):
                    os.remove(path)
    return TARGET


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
