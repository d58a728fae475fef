"""In-tree builds that are safe to trigger from several processes and on copied trees.

A build is redone when the output is missing or when the SHA-256 of its dependencies (contents,
not mtimes: the tree travels to other machines whose clocks may differ) is not the one recorded
in `<output>.stamp`. A file lock serialises concurrent callers (one process per GPU under
torchrun) and the output is written to a temporary file and renamed into place, so a process that
already mapped the previous library never sees it change underneath."""
import fcntl
import hashlib
import os
import subprocess
import sys


def fingerprint(deps, extra=""):
    h = hashlib.sha256(extra.encode())
    for d in deps:
        h.update(os.path.basename(d).encode() + b"\0")
        with open(d, "rb") as f:
            h.update(f.read())
        h.update(b"\0")
    return h.hexdigest()


def is_current(out, deps, extra=""):
    stamp = out + ".stamp"
    if not (os.path.exists(out) and os.path.exists(stamp)):
        return False
    try:
        with open(stamp) as f:
            return f.read().strip() == fingerprint(deps, extra)
    except OSError:
        return False


def ensure_built(out, deps, make_cmd, extra="", force=False, verbose=False):
    """`make_cmd(tmp_out)` returns the argv that writes `tmp_out`. Returns `out`."""
    deps = [d for d in deps if os.path.exists(d)]
    if not force and is_current(out, deps, extra):
        return out
    with open(out + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current(out, deps, extra):  # another process built it meanwhile
                return out
            tmp = f"{out}.tmp.{os.getpid()}"
            res = subprocess.run(make_cmd(tmp), capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.unlink(tmp)
                raise RuntimeError("build failed: " + out)
            os.replace(tmp, out)
            with open(out + ".stamp.tmp", "w") as f:
                f.write(fingerprint(deps, extra) + "\n")
            os.replace(out + ".stamp.tmp", out + ".stamp")
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return out
