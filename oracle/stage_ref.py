"""Stage the UNMODIFIED reference for the GPU box (test infrastructure; recipe only — no reference source is committed).

    python oracle/stage_ref.py        # also called by __graft_entry__.build() when /root/reference exists

/root/reference does not exist on the GPU box, so parity tests there could only use committed golden vectors. This copies the
pure-Python packages the hot path touches — /root/reference/tracklab (engine, pipeline, datastruct, wrappers) and
/root/reference/plugins/track (the tracker plugins) — verbatim into the git-ignored directory oracle/_ref/, which travels to the
box with the snapshot like the built .so files do. oracle/ref_env.install() then imports the reference from there, so
tests/test_real_engine_gpu.py drives the REAL OfflineTrackingEngine + TrackerState over the CUDA modules and bench.py's
reference arm times the REAL StrongSORT plugin. Nothing under oracle/_ref/ is tracked by git."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
PARTS = (("tracklab", "tracklab"), (os.path.join("plugins", "track"), os.path.join("plugins", "track")))


def stage(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("oracle/stage_ref.py: /root/reference absent, nothing staged (the GPU box uses what the build container staged)")
        return False
    for a, b in PARTS:
        dst = os.path.join(DST, b)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, a), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.pth", "*.pt", "*.onnx"))
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(SRC + "\n")
    if verbose:
        print(f"oracle/stage_ref.py: staged {[b for _, b in PARTS]} into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() or True else 1)
