"""SASS evidence of the Blackwell-native instructions in libvibo_b200.so (no GPU needed):

    python profiles/sass_summary.py > profiles/r02_sass_tcgen05.md

Counts, per kernel, the mnemonics B200_PROFILING.md names: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM,
tcgen05.commit -> UTCBAR, tcgen05.alloc -> UTCATOMSWS, 2-D TMA -> UTMALDG, 1-D bulk copy -> UBLKCP,
mma.sync -> HMMA, packed f32x2 arithmetic -> FFMA2 / FADD2 / FMUL2.
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "variational-item-response-theory-public_b200", "libvibo_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UBLKCP", "HMMA", "FFMA2", "FADD2", "FMUL2"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = name.replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "")
            per.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            for k in KEYS:
                if op == k or (k.startswith("UTC") and op.startswith(k)):
                    per[cur][k] += 1
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print("# SASS evidence: Blackwell-native instructions in libvibo_b200.so (round 2)\n")
    print("`python profiles/sass_summary.py` (`cuobjdump -sass` of the shipped library, mnemonics counted per kernel; "
          "B200_PROFILING.md: `tcgen05.mma` -> `UTC*MMA`, `tcgen05.ld` -> `LDTM`, `tcgen05.commit` -> `UTCBAR`, "
          "`tcgen05.alloc` -> `UTCATOMSWS`, 2-D TMA -> `UTMALDG`, 1-D bulk copy -> `UBLKCP`, `mma.sync` -> `HMMA`, "
          "packed f32x2 arithmetic -> `FFMA2 / FADD2 / FMUL2`).\n")
    print("## whole library\n\n| mnemonic | count |\n|---|---:|")
    for k in KEYS:
        if total[k]:
            print(f"| `{k}` | {total[k]} |")
    print("\n## kernels that issue tcgen05 / TMEM / 2-D TMA instructions (all template instantiations summed)\n")
    print("| kernel | UTCHMMA | LDTM | UTCBAR | UTCATOMSWS | UTMALDG | FFMA2 |\n|---|---:|---:|---:|---:|---:|---:|")
    fam = collections.OrderedDict()
    for name, c in per.items():
        base = name.split("<")[0]
        fam.setdefault(base, collections.Counter()).update(c)
        fam[base]["_n"] += 1
    for base, c in fam.items():
        if c["UTCHMMA"] or c["LDTM"] or c["UTMALDG"]:
            print(f"| `{base}` ({c['_n']} instantiation{'s' if c['_n'] > 1 else ''}) | {c['UTCHMMA']} | {c['LDTM']} | "
                  f"{c['UTCBAR']} | {c['UTCATOMSWS']} | {c['UTMALDG']} | {c['FFMA2']} |")
    print("\n## kernels that stream rows with 1-D TMA bulk copies (`UBLKCP`) and packed f32x2 arithmetic\n")
    print("| kernel | instantiations | UBLKCP | FFMA2 | FADD2 | FMUL2 | HMMA |\n|---|---:|---:|---:|---:|---:|---:|")
    for base, c in fam.items():
        if c["UBLKCP"]:
            print(f"| `{base}` | {c['_n']} | {c['UBLKCP']} | {c['FFMA2']} | {c['FADD2']} | {c['FMUL2']} | {c['HMMA']} |")


if __name__ == "__main__":
    main()
