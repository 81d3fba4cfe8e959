"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS.*,
plus legacy-path markers that must NOT appear (HMMA = mma.sync, BRA.U.ANY next to UTCHMMA = the per-MMA elect loop
ptxas emits when the issuing thread is chosen with a threadIdx predicate instead of elect.sync).

    python tools/sass_mnemonics.py > profiles/sass_mnemonics.txt      # build container: needs cuobjdump, no GPU
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"\b(UTC[A-Z0-9]*MMA[A-Z0-9.]*|UTMALDG[A-Z0-9.]*|UTMASTG[A-Z0-9.]*|UBLKCP[A-Z0-9.]*|LDTM[A-Z0-9.]*|"
                 r"STTM[A-Z0-9.]*|UTCBAR[A-Z0-9.]*|UTCATOMSWS[A-Z0-9.]*|SYNCS\.[A-Z0-9.]*|ELECT|HMMA[A-Z0-9.]*|"
                 r"BRA\.U\.ANY|MUFU\.EX2|REDG[A-Z0-9.]*|LDG\.E[A-Z0-9.]*\.CONSTANT|ST\.E[A-Z0-9.]*SYS[A-Z0-9.]*|LD\.E[A-Z0-9.]*SYS[A-Z0-9.]*)")


def main():
    lib = os.path.join(ROOT, "moco_b200", "libmoco_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.splitlines()
    counts, order, cur, i = collections.defaultdict(collections.Counter), [], None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", names[i]).replace("moco::", "") if i < len(names) else m.group(1)
            i += 1
            order.append(cur)
            continue
        if cur and "/*" in line:
            for mn in PAT.findall(line.split("/*")[1] if line.strip().startswith("/*") else line):
                counts[cur][mn] += 1
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print(f"# SASS mnemonics per kernel of moco_b200/libmoco_b200.so ({', '.join(sorted(set(arch)))}); static instruction counts")
    print("# tcgen05.mma -> UTCHMMA[.2CTA], tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP,")
    print("# tcgen05.commit -> UTCBAR, mbarrier -> SYNCS.*, elect.sync -> ELECT; HMMA (legacy mma.sync) and BRA.U.ANY")
    print("# (per-instruction elect loops around uniform-datapath ops) should be absent from the tensor-core kernels")
    for fn in order:
        c = counts[fn]
        if not c:
            continue
        print(f"\n{fn}")
        for mn, n in sorted(c.items()):
            print(f"    {n:5d}  {mn}")
    bad = {fn: c for fn, c in counts.items() if c.get("HMMA") or (c.get("BRA.U.ANY") and any(k.startswith("UTCHMMA") for k in c))}
    print("\n# legacy / elect-loop markers in tensor-core kernels:", "none" if not bad else dict((k, dict(v)) for k, v in bad.items()))


if __name__ == "__main__":
    main()
