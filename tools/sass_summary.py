"""Per-kernel SASS opcode summary of libdae.so (what proves TMA / mbarrier / cluster / cp.async use; see
/opt/skills/guides/B200_PROFILING.md "What proves a Blackwell-native kernel").

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dynamic-asr-eval_b200", "libdae.so")
WATCH = ["CREDUX", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "UCGABAR", "MUFU.EX2", "MUFU.LG2", "MUFU.RCP", "SHFL", "REDUX",
         "BAR.SYNC", "LDS", "STS", "LDG", "STG", "ATOM", "RED", "DADD", "DFMA", "DMUL", "FFMA", "FADD", "FMUL", "FMNMX",
         "HMMA", "UTC"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern, counts, total = None, {}, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", kern).replace("dae::", "")
            counts[kern], total[kern] = collections.Counter(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if kern and m:
            op = m.group(1)
            total[kern] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w.endswith(".") is False and op.startswith(w) and w in ("UTC", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "UCGABAR", "SHFL", "REDUX", "HMMA")):
                    counts[kern][w] += 1
                    break
    print("# SASS opcode summary of libdae.so (sm_100a), static instruction counts per kernel\n")
    print("`UBLKCP` = TMA bulk copy (cp.async.bulk), `SYNCS` = mbarrier, `UCGABAR` = cluster barrier, `LDGSTS` = cp.async,")
    print("`REDUX` = warp reduce.  None of these kernels is a contraction, so no `UTC*MMA` / `HMMA` is expected.\n")
    cols = [w for w in WATCH if any(counts[k][w] for k in counts)]
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for k in sorted(counts):
        print(f"| `{k}` | {total[k]} | " + " | ".join(str(counts[k][w]) if counts[k][w] else "" for w in cols) + " |")


if __name__ == "__main__":
    main()
