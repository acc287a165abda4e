"""Static view of one kernel's SASS: instructions per source line and per loop (backward branch), from `nvdisasm -g -c`.
usage: python tools/sass_loops.py <object.o> <kernel-substring> [--lines]
Used to estimate warp-instruction counts of a kernel variant without a GPU (trip counts come from the model dimensions)."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def disasm(obj, sub):
  tmp = tempfile.mkdtemp()
  subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
  cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
  txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
  starts = [i for i, l in enumerate(txt) if l.startswith("\t.section\t.text.")]
  for k, s in enumerate(starts):
    if sub in txt[s]:
      return txt[s:(starts[k + 1] if k + 1 < len(starts) else len(txt))]
  raise SystemExit("kernel not found: " + sub)


def parse(lines):
  ins, labels, cur = [], {}, ("?", 0)
  for l in lines:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
      cur = (os.path.basename(m.group(1)), int(m.group(2)))
      continue
    m = re.match(r"^(\.L_x_\d+):", l)
    if m:
      labels[m.group(1)] = len(ins)
      continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
      ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
  return ins, labels


def main():
  obj, sub = sys.argv[1], sys.argv[2]
  ins, labels = parse(disasm(obj, sub))
  print("instructions:", len(ins))
  loops = []
  for i, (_, text, _) in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)*\s+.*`\((\.L_x_\d+)\)", text)
    if m and m.group(1) in labels and labels[m.group(1)] <= i:
      loops.append((labels[m.group(1)], i))
  print("loops (start idx, size, dominant source lines):")
  for s, e in sorted(loops):
    c = collections.Counter(x[2] for x in ins[s:e + 1])
    ops = collections.Counter(x[1].split()[0].lstrip("@!P0123456789UT ").split(".")[0] if not x[1].startswith("@") else x[1].split()[1].split(".")[0] for x in ins[s:e + 1])
    top = ", ".join(f"{f}:{ln}x{n}" for (f, ln), n in c.most_common(4))
    print(f"  [{s:5d}..{e:5d}] n={e - s + 1:4d}  {top}   | " + " ".join(f"{k}{v}" for k, v in ops.most_common(6)))
  if "--lines" in sys.argv:
    c = collections.Counter(x[2] for x in ins)
    for (f, ln), n in sorted(c.items()):
      print(f"{f}:{ln}\t{n}")


main()
