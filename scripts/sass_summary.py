"""Per-kernel SASS instruction mix of the product library (cuobjdump -sass): loads / stores by width, local-memory traffic (spills),
shared memory, atomics, shuffles, barriers, fp64 arithmetic.  Writes profiles/<round>_sass_summary.txt."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_summary.txt")
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "lrr_b200", "liblrr_b200.so")], capture_output=True, text=True).stdout
kern, rows = None, collections.OrderedDict()
GROUPS = [("LDG.E.128", r"\bLDG\.E\.128"), ("LDG.E.64", r"\bLDG\.E\.64"), ("LDG other", r"\bLDG\b"), ("STG", r"\bSTG\b"), ("LDL (spill load)", r"\bLDL\b"),
          ("STL (spill store)", r"\bSTL\b"), ("LDS", r"\bLDS\b"), ("STS", r"\bSTS\b"), ("ATOM/RED", r"\b(ATOM|ATOMG|RED|ATOMS)\b"), ("SHFL", r"\bSHFL\b"),
          ("BAR", r"\bBAR\b"), ("DFMA", r"\bDFMA\b"), ("DADD", r"\bDADD\b"), ("DMUL", r"\bDMUL\b"), ("MUFU", r"\bMUFU\b"), ("LDGSTS/UBLKCP/UTMA", r"\b(LDGSTS|UBLKCP|UTMALDG)\b")]
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kern = name; rows[kern] = collections.Counter(); continue
    if kern and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        ins = line.split("*/", 1)[1]
        rows[kern]["total"] += 1
        matched_ldg = False
        for g, pat in GROUPS:
            if re.search(pat, ins):
                if g.startswith("LDG"):
                    if matched_ldg: continue
                    matched_ldg = True
                rows[kern][g] += 1
with open(out, "w") as f:
    f.write("SASS instruction mix per kernel (static counts), liblrr_b200.so, sm_100a; cuobjdump -sass | scripts/sass_summary.py\n")
    f.write("No tensor-core (UTCMMA/HMMA) or TMA (UTMALDG/UBLKCP) instructions: the path is fp64 gather + reduction (DESIGN.md §2).\n\n")
    cols = ["total"] + [g for g, _ in GROUPS]
    f.write(f"{'kernel':28s}" + "".join(f"{c[:12]:>13s}" for c in cols) + "\n")
    for k, c in rows.items():
        f.write(f"{k[:28]:28s}" + "".join(f"{c.get(col, 0):13d}" for col in cols) + "\n")
print(open(out).read()[:3000])
