"""Warp-stall breakdown of every kernel in an `ncu --set full` report (run HERE, on the CPU: `ncu -i` only reads the file):
the PC-sampling counters `smsp__pcsamp_warps_issue_stalled_*` as shares of all samples, plus the source lines that
collected the most samples (`--page source`, needs `-lineinfo` + `--import-source on`).
  python scripts/ncu_stalls.py gpurun_out/r02final/top_kernels_c2.ncu-rep > profiles/r02_stalls_find_insert.txt"""
import csv
import io
import re
import subprocess
import sys


def raw(rep):
  out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  r = list(csv.reader(io.StringIO(out)))
  return r[0], r[2:]


def source(rep, kernel_id):
  out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", "::%s:" % kernel_id],
                       capture_output=True, text=True).stdout
  return out


def hot_sass(rep, top=8):
  """kernel name -> [(samples, share, SASS text)] from the SASS view of the source page"""
  out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
  res, name, hdr, acc = {}, None, None, []
  for row in csv.reader(io.StringIO(out)):
    if not row:
      continue
    if row[0] == "Kernel Name":
      if name is not None:
        res[name] = acc
      name, hdr, acc = re.sub(r"\(.*", "", row[1]).replace("void ", "").replace("det::", "").replace("(int)", ""), None, []
    elif row[0] == "Address":
      hdr = row
    elif hdr is not None and name is not None:
      try:
        acc.append((int(row[hdr.index("# Samples")]), row[hdr.index("Source")].strip()))
      except (ValueError, IndexError):
        pass
  if name is not None:
    res[name] = acc
  outd = {}
  for k, v in res.items():
    tot = sum(n for n, _ in v) or 1
    outd[k] = [(n, 100.0 * n / tot, t) for n, t in sorted(v, reverse=True)[:top]]
  return outd


def main():
  rep = sys.argv[1]
  sass = hot_sass(rep)
  h, rows = raw(rep)
  ki = h.index("Kernel Name")
  cols = [i for i, c in enumerate(h) if c.startswith("smsp__pcsamp_warps_issue_stalled_") and not c.endswith("_not_issued")]
  for row in rows:
    name = re.sub(r"\(.*", "", row[ki]).replace("void ", "")
    vals = []
    for i in cols:
      try:
        vals.append((float(row[i]), h[i].replace("smsp__pcsamp_warps_issue_stalled_", "")))
      except ValueError:
        pass
    total = sum(v for v, _ in vals) or 1.0
    vals.sort(reverse=True)
    print("%s: %d samples" % (name, int(total)))
    for v, c in vals[:6]:
      print("   %5.1f %%  %s" % (100.0 * v / total, c))
    key = name.replace("det::", "").replace("(int)", "")
    for k2, lines in sass.items():
      if key.strip().startswith(k2.split("<")[0].strip()):
        print("   instructions with the most samples (the sample lands on the instruction that WAITS):")
        for n, share, text in lines:
          print("     %5.1f %%  %s" % (share, text))
    print()


if __name__ == "__main__":
  main()
