"""Count the Blackwell-specific SASS mnemonics in the built objects (evidence for profiles/): UTCHMMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce), UBLKCP (bulk copy), SYNCS (mbarrier)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ccnet_b200", "lib")
PAT = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UBLKCP|UTCBAR|UTCATOMSWS|SYNCS|HMMA|FFMA|MUFU\.EX2)\b")


def main():
    objs = sorted(f for f in os.listdir(LIB) if f.endswith(".o"))
    print("cuobjdump -sass of ccnet_b200/lib/*.o (sm_100a), mnemonic counts per kernel\n")
    for o in objs:
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(LIB, o)], capture_output=True, text=True).stdout
        cur, counts = None, collections.OrderedDict()
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                cur = cur.replace("cca::(anonymous namespace)::", "").replace("(anonymous namespace)::", "")
                cur = re.sub(r"\(CUtensorMap.*", "", cur)
                counts[cur] = collections.Counter()
                continue
            if cur:
                for mm in PAT.findall(line):
                    counts[cur][mm.split(".")[0] if not mm.startswith("MUFU") else mm] += 1
        if not counts:
            continue
        print(f"== {o}")
        for k, c in counts.items():
            if any(n.startswith(("UTC", "LDTM", "STTM", "UTMA", "UBLKCP")) for n in c):
                print(f"  {k}")
                print("     " + "  ".join(f"{n}={v}" for n, v in sorted(c.items())))
    return 0


if __name__ == "__main__":
    sys.exit(main())
