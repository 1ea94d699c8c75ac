"""Per-kernel resource table from the ptxas logs of the last build (csrc/Makefile passes -Xptxas -v and keeps
build/<file>.ptxas.log): registers per thread, static shared memory, stack frame, spill bytes, barriers.  No GPU needed.

  python tools/ptxas_summary.py small-vision_b200/csrc/build > profiles/rNN_ptxas_summary.txt

Dynamic shared memory (the TMA rings of the GEMM / attention kernels, set with cudaFuncSetAttribute at launch) is not known to
ptxas; the launch code states it (gemm.cu, attention_tc.cu, elementwise.cu).
"""
import glob
import os
import re
import subprocess
import sys


def demangle(names):
  r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True)
  return r.stdout.splitlines()


def short(name):
  name = re.sub(r"\(anonymous namespace\)::", "", name)
  name = re.sub(r"^void ", "", name)
  name = re.sub(r"\(.*\)$", "", name)          # drop the argument list, keep template arguments
  return name


def main(build_dir):
  rows = []
  for path in sorted(glob.glob(os.path.join(build_dir, "*.ptxas.log"))):
    unit = os.path.basename(path)[:-len(".ptxas.log")]
    cur = None
    for line in open(path):
      m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
      if m:
        cur = {"unit": unit, "name": m.group(1), "stack": 0, "spill_st": 0, "spill_ld": 0, "regs": 0, "smem": 0, "bar": 0}
        rows.append(cur)
        continue
      if cur is None:
        continue
      m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
      if m:
        cur["stack"], cur["spill_st"], cur["spill_ld"] = (int(x) for x in m.groups())
      m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?", line)
      if m:
        cur["regs"] = int(m.group(1))
        cur["bar"] = int(m.group(2) or 0)
        s = re.search(r"(\d+) bytes smem", line)
        cur["smem"] = int(s.group(1)) if s else 0
  names = demangle([r["name"] for r in rows])
  for r, n in zip(rows, names):
    r["short"] = short(n)
  print(f"# ptxas -v over {len(rows)} kernels (sm_100a), from {build_dir}/*.ptxas.log")
  spilled = [r for r in rows if r["spill_st"] or r["spill_ld"]]
  fam = {}
  for r in spilled:
    base = r["short"].split("<")[0]
    n, st, ld = fam.get(base, (0, 0, 0))
    fam[base] = (n + 1, max(st, r["spill_st"]), max(ld, r["spill_ld"]))
  print(f"# kernels with register spills: {len(spilled)} of {len(rows)}" + ("" if not spilled else " — " + "; ".join(
      f"{b}: {n} instantiation(s), at most {st} B stored / {ld} B loaded per thread" for b, (n, st, ld) in fam.items())))
  print(f"# max registers per thread: {max(r['regs'] for r in rows)}; kernels with a stack frame: {sum(1 for r in rows if r['stack'])}")
  print(f"{'file':<16}{'regs':>5}{'smem(static)':>14}{'stack':>7}{'spill st/ld':>13}{'bar':>5}  kernel")
  for r in rows:
    print(f"{r['unit']:<16}{r['regs']:>5}{r['smem']:>14}{r['stack']:>7}{str(r['spill_st']) + '/' + str(r['spill_ld']):>13}{r['bar']:>5}  {r['short']}")


if __name__ == "__main__":
  main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "small-vision_b200", "csrc", "build"))
