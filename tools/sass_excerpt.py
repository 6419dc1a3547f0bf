#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell path (tcgen05 / TMEM / TMA bulk
copies / mbarriers / setmaxnreg / DSMEM stores) in the built library; writes profiles/r02_sass_excerpt.txt.

  python tools/sass_excerpt.py [graphcast_b200/libgraphcast_b200.so] > profiles/r02_sass_excerpt.txt
"""
import collections
import re
import subprocess
import sys

PAT = re.compile(r"\b(UTCHMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UBLKCP|UBLKPF|UTMALDG|SYNCS|USETMAXREG|STAS|ELECT|"
                 r"UCGABAR_ARV|UCGABAR_WAIT|FENCE\.VIEW\.ASYNC|MEMBAR\.ALL)[\w.]*")


def main():
  lib = sys.argv[1] if len(sys.argv) > 1 else "graphcast_b200/libgraphcast_b200.so"
  sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
  names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)),
                         capture_output=True, text=True).stdout.splitlines()
  counts, order, cur = {}, [], None
  it = iter(names)
  for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
      cur = next(it, m.group(1))
      counts[cur] = collections.Counter()
      order.append(cur)
      continue
    if cur is None:
      continue
    for m in PAT.finditer(line.split("/*")[1] if "/*" in line else line):
      counts[cur][m.group(0)] += 1
  print(f"# SASS evidence of the Blackwell path in {lib} (final round-2 build)")
  print("# tools/sass_excerpt.py: cuobjdump -sass | per kernel: mnemonic x count")
  print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld (TMEM -> registers), "
        "UBLKCP = cp.async.bulk (TMA engine),")
  print("# UBLKPF = cp.async.bulk.prefetch.L2, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, "
        "STAS = st.async (DSMEM), UTCATOMSWS = tcgen05.alloc")
  for k in sorted(order):
    print()
    print(k)
    print("   " + ("  ".join(f"{n} x{c}" for n, c in sorted(counts[k].items())) or "(none: plain SIMT kernel)"))


if __name__ == "__main__":
  main()
