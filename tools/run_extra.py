"""Run one of bench.py's extra workloads alone and print its JSON object.
Usage: python tools/run_extra.py c3 | c5 [scans] | scan_loop      (IKD_PHASES=1 adds the library's phase / allocation trace on stderr)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    which = sys.argv[1]
    sys.argv = [sys.argv[0]] + [a for a in sys.argv[2:] if a.startswith("--")]
    args = bench.parse()
    rest = [a for a in sys.argv[1:] if not a.startswith("--")]
    if which == "c3":
        out = bench.c3_extra(args, 0)
    elif which == "c5":
        scans = int(os.environ.get("C5_SCANS", "1000"))
        out = bench.c5_extra(args, 0, scans=scans)
    elif which == "scan_loop":
        out = bench.scanloop_ours(args, 0, 1, 0, 3, 20, with_plane=False)
    else:
        raise SystemExit(__doc__)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
