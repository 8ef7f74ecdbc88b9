"""Join an ncu SASS source-page CSV with nvdisasm line info -> per source line instruction / stall-sample totals.
usage: ncu_lines.py <ncu-rep> <libb200sim.so> <kernel regex> [topN]"""
import csv, re, subprocess, sys, os, tempfile, collections
rep, so, kre = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sel = sys.argv[5] if len(sys.argv) > 5 else "CapsILi12ELi4ELi32ELi1EEELi9E"  # mangled-name fragment choosing the instantiation
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(so)} >/dev/null", shell=True, check=True)
dis = []
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    dis += subprocess.run(f"nvdisasm --print-line-info -c {tmp}/{cubin}", shell=True, capture_output=True, text=True).stdout.splitlines()
# map offset -> (file,line) for the first section whose name matches
line_of = {}
cur = None; insec = False
for l in dis:
    if l.startswith("\t.section\t.text."):
        insec = re.search(kre, l) is not None and (sel in l)
        cur = None
    if not insec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/", l)
    if m: line_of[int(m.group(1), 16)] = cur
out = subprocess.run(f"ncu -i {rep} --page source --csv --kernel-name regex:{kre}", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hdr_i]
ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0]
for r in rows[hdr_i + 1:]:
    if len(r) <= iinst or not r[ia].startswith("0x"): 
        if r and r[0] == "Kernel Name": break
        continue
    a = int(r[ia], 16)
    if base is None: base = a
    key = line_of.get(a - base)
    s = int(r[isamp] or 0); n = int(r[iinst] or 0)
    agg[key][0] += s; agg[key][1] += n; agg[key][2] += 1
    tot[0] += s; tot[1] += n
print("total samples", tot[0], "warp-instructions", tot[1], "static SASS", sum(v[2] for v in agg.values()))
byfile = collections.defaultdict(lambda: [0, 0, 0])
for k, v in agg.items():
    f = k[0] if k else None
    for i in range(3): byfile[f][i] += v[i]
for f, v in sorted(byfile.items(), key=lambda x: -x[1][0]): print(f"  {f}: samples {v[0]} ({100*v[0]/max(tot[0],1):.1f}%) inst {v[1]} ({100*v[1]/max(tot[1],1):.1f}%) sass {v[2]}")
print("top lines by samples:")
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"  {k}: samples {v[0]} ({100*v[0]/max(tot[0],1):.1f}%) inst {v[1]} ({100*v[1]/max(tot[1],1):.1f}%) sass {v[2]}")
