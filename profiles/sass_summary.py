#!/usr/bin/env python
"""SASS evidence of the Blackwell-native paths in spex_b200/libspex_b200.so (no GPU needed):
    python profiles/sass_summary.py > profiles/r02_sass_mnemonics.txt
Counts, per kernel, the mnemonics B200_PROFILING.md lists: UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st),
UBLKCP/UTMALDG (bulk / tensor async copy), UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (legacy mma.sync:
must be absent).  multimem.st.relaxed.sys.global.v4.f32 (csrc/common.cuh: st_multimem_f4) has no mnemonic of its
own: it is emitted as STG.E.128.STRONG.SYS on the NVSwitch multicast address, so those stores are counted too
(kernels: the SpMM epilogues, mcast_rows_kernel); P2P peer stores are plain STG.E.128.  LDG.E.NA.{EL,EF,EN}L2.256 are
the SpMM's 256-bit gathers with a static L2 eviction priority (evict_last / evict_first / normal)."""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "spex_b200/libspex_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA|LDTM\S*|STTM\S*|UBLKCP\S*|UTMALDG\S*|UTCBAR\S*|UTCATOM\S*|SYNCS\S*|HMMA\S*|HGMMA\S*|"
                 r"VHMNMX|HMNMX2|FMNMX3?|ELECT|REDUX\S*|STG\.E\.128\.STRONG\.SYS|LDGSTS\S*|LDG\.E\.NA\.E[LFN]L2\.256\S*)")
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for tok in pat.findall(line):
        per[cur][tok.split(".")[0] if not tok.startswith(("LDTM", "UBLKCP", "STG", "LDG")) else tok] += 1
print(f"# cuobjdump -sass {LIB} : Blackwell mnemonics per kernel (kernels without any are omitted)")
tot = collections.Counter()
for k, c in per.items():
    if c:
        print(f"{k[:110]}")
        print("    " + "  ".join(f"{a}x{b}" for a, b in sorted(c.items())))
        tot.update(c)
print("# totals: " + "  ".join(f"{a}x{b}" for a, b in sorted(tot.items())))
print("# legacy tensor path (HMMA / HGMMA) present:", any(a.startswith(("HMMA", "HGMMA")) for a in tot))
