"""Per-kernel counts of the SASS instructions that prove the Blackwell-native path (B200_PROFILING.md): UTCHMMA / UTCQMMA
(tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAPF (TMA load / store / prefetch), UTCBAR
(tcgen05.commit), SYNCS (mbarrier), REDG / RED (reductions to global), MUFU.  No GPU needed.

  python tools/sass_summary.py small-vision_b200/libumd_b200.so > profiles/rNN_sass_summary.txt
"""
import collections
import re
import subprocess
import sys

PAT = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "RED", "MUFU", "HMMA", "FFMA"]


def main(path):
  out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
  demangle = {}
  cur = None
  counts = collections.OrderedDict()
  for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
      cur = m.group(1)
      counts[cur] = collections.Counter()
      continue
    if cur is None:
      continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
      continue
    op = m.group(1)
    counts[cur]["_total"] += 1
    for p in PAT:
      if op.startswith(p):
        counts[cur][p] += 1
        break
  names = list(counts)
  try:
    dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, dm))
  except Exception:
    pass
  head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
  print(f"# cuobjdump -sass {path} (sm_100a), commit {head}: instruction counts per kernel")
  print(f"# {'kernel':88s} " + " ".join(f"{p:>8s}" for p in PAT) + f" {'total':>8s}")
  tot = collections.Counter()
  for k, c in counts.items():
    name = re.sub(r"\(.*$", "", demangle.get(k, k)).replace("void ", "").replace("umd::", "").replace("(anonymous namespace)::", "")
    print(f"{name[:90]:90s} " + " ".join(f"{c[p]:8d}" for p in PAT) + f" {c['_total']:8d}")
    tot.update(c)
  print(f"{'ALL KERNELS':90s} " + " ".join(f"{tot[p]:8d}" for p in PAT) + f" {tot['_total']:8d}")


if __name__ == "__main__":
  main(sys.argv[1] if len(sys.argv) > 1 else "small-vision_b200/libumd_b200.so")
