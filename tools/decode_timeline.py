"""Decode the TCTL lines a -DINRF_TC_TIMELINE build of libinrf prints (tools/gpu_timeline.sh).

    python tools/decode_timeline.py gpurun_out/timeline_fused.log [S=64|192|any] [roles=0,1,3] [max_cycle]

One line per role: `event@cycle(+delta to the previous printed event)`; cycles are relative to the earliest stamp of the
launch.  Roles: 0 issuer (thread 480), 1 / 2 epilogue warps 0 / 4 (threads 0 / 128; their per-chunk ld / cvt / arrive stamps
are folded away unless roles contains 'x'), 3 back-end warp 12 (thread 384)."""
import re
import sys

ISSUER = {1: "fenced", 2: "issued", 3: "tile", 4: "F_READY", 5: "TAIL_DONE", 6: "V_READY", 7: "A_RDY(a2s2)", 13: "acq", 14: "pre", 15: "probed"}
EPILOGUE = {0: "tail_done", 3: "tile", 6: "rows_out", 7: "ring_free", 8: "ld", 9: "cvt", 10: "arrive", 11: "FIRST", 12: "FULL", 13: "VIEWS",
            14: "ALBSH", 15: "SMALL"}
BACKEND = {3: "be_tile", 1: "RAW_READY", 2: "seg", 5: "ray_end", 6: "resampled", 4: "RAW_FREE"}


def main():
    path = sys.argv[1]
    want_s = sys.argv[2] if len(sys.argv) > 2 else "any"
    roles_arg = sys.argv[3] if len(sys.argv) > 3 else "0,1,3"
    lim = int(sys.argv[4]) if len(sys.argv) > 4 else 10 ** 12
    verbose = "x" in roles_arg
    roles = [int(r) for r in roles_arg.replace("x", "").split(",") if r]
    lines = [l for l in open(path) if l.startswith("TCTL") and " role=" in l]
    if want_s != "any":
        lines = [l for l in lines if f" S={want_s} " in l]
    for l in lines[-4:]:                      # the last launch that matches
        m = re.match(r"TCTL (?:fuse=(\d) S=(\d+) )?role=(\d) n=(\d+):(.*)", l)
        role = int(m.group(3))
        if role not in roles:
            continue
        names = ISSUER if role == 0 else (BACKEND if role == 3 else EPILOGUE)
        ev = [(int(a), int(b)) for a, b in (x.split(":") for x in m.group(5).split())]
        print(f"fuse={m.group(1)} S={m.group(2)} role={role} stamps={len(ev)}")
        prev, out = None, []
        for t, k in ev:
            if t > lim:
                break
            if role in (1, 2) and not verbose and names.get(k) in ("ld", "cvt", "arrive"):
                prev = t
                continue
            out.append(f"{names.get(k, k)}@{t}" + ("" if prev is None else f"(+{t - prev})"))
            prev = t
        print(" ".join(out))


if __name__ == "__main__":
    main()
