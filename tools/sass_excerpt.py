"""cuobjdump -sass of one kernel of libb200sim.so -> header (instruction count, opcode histogram) + the first N instructions.
usage: sass_excerpt.py <mangled-name fragment> <title> [N]"""
import collections, re, subprocess, sys, os
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "maniskill_b200", "libb200sim.so")
frag, title = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 120
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.splitlines()
cur, body = None, {}
for l in out:
    m = re.match(r"\s+Function : (\S+)", l)
    if m:
        cur = m.group(1); body[cur] = []; continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        body[cur].append(l.rstrip())
names = [k for k in body if frag in k]
print(f"cuobjdump -sass -- {title}\n")
for k in names:
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", l.split("*/", 1)[1].strip()).split()[0].split(".")[0].rstrip(";") for l in body[k])
    print(f"Function : {k}   ({len(body[k])} SASS instructions)")
    print("opcode histogram: " + ", ".join(f"{o} {c}" for o, c in ops.most_common(24)))
    print("\n".join(body[k][:n]))
    print()
