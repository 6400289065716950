"""profiles/r02_sass_summary.txt: per kernel of every sm_100a object under geoformer_b200/build, the instruction count
and the counts of the mnemonics that characterise it (atomics, barriers, vector accesses, cluster / TMA / tcgen05
instructions).  Needs only cuobjdump (no GPU):  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"(REDG|RED\.|ATOM|BAR|LDG\.E\.128|STG\.E\.(EF\.)?128|LDS|STS|SHFL|VOTE|REDUX|CREDUX|UCGABAR|STAS|SYNCS|"
                 r"CCTL|MEMBAR|FLO|POPC|MUFU|LDSM|UTMA|UTC|LDTM|STTM|UTCBAR|R2UR|NANOSLEEP|ELECT|UBLKCP|UTCHMMA|UTCQMMA)")


def main():
    print("# SASS evidence: cuobjdump -sass of the sm_100a objects in geoformer_b200/build (python -m geoformer_b200.build)")
    print("# per kernel: instruction count, then the characteristic mnemonics with their counts\n")
    for path in sorted(glob.glob(os.path.join(ROOT, "geoformer_b200", "build", "*.o"))):
        txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
        cur, stats = None, collections.OrderedDict()
        for line in txt.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                stats[cur] = collections.Counter()
                continue
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and cur:
                stats[cur][m.group(1)] += 1
        print("## %s  (%s)\n" % (os.path.basename(path), ",".join(arch)))
        for fn, c in stats.items():
            dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
            agg = collections.Counter()
            for k, v in c.items():
                if PAT.match(k):
                    parts = k.split(".")
                    agg[".".join(parts[:3]) if parts[0] in ("LDG", "STG") else ".".join(parts[:2])] += v
            print("%s\n    %d instructions; %s\n" % (dem[:160], sum(c.values()),
                                                     ", ".join("%s x%d" % kv for kv in sorted(agg.items(), key=lambda kv: -kv[1])[:18])))


if __name__ == "__main__":
    main()
