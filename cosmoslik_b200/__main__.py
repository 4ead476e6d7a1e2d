"""``python -m cosmoslik_b200 script.py [script args]`` -- the minimal launcher (reference:
cosmoslik/__main__.py:44-45: load the script, build Slik, drain the sampler).  The reference's
``-n N`` (mpiexec re-exec) and qsub options are replaced by ``torchrun --nproc-per-node N -m
cosmoslik_b200 script.py``: ranks shard the walkers / chains."""
import argparse
import sys

from . import dist
from .core import Slik, load_script


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m cosmoslik_b200")
    ap.add_argument("script")
    ap.add_argument("--traceback", action="store_true")
    args, rest = ap.parse_known_args(argv)
    dist.init_from_env()
    parser, script = load_script(args.script)
    kw = vars(parser.parse_args(rest))
    try:
        for _ in Slik(script(**kw)).sample():
            pass
    except Exception:
        if args.traceback:
            raise
        print("error: %s" % (sys.exc_info()[1],), file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
