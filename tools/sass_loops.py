"""Development tool: per-loop SASS opcode histogram of one kernel in a cubin / .so (no GPU needed).

    python tools/sass_loops.py cardiax_b200/csrc/libfk.so 'fk_stream_kernelILb0ELi2ELb1ELb0E' [--min 100]

Finds backward branches (loops), prints for each loop body its instruction count and opcode histogram -- the steady-state
loop of the streaming kernel is the largest one.  Used to count thread-instructions per cell-step before going to ncu.
"""
import collections
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, res = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and cur:
            res[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return res


def opcode(text):
    t = text.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0]


def main():
    path, pat = sys.argv[1], sys.argv[2]
    minlen = int(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 100
    for name, ins in kernels(path).items():
        if pat not in name:
            continue
        print("== %s: %d instructions" % (name, len(ins)))
        loops = []
        for addr, text in ins:
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
            if m and int(m.group(1), 16) < addr:
                loops.append((int(m.group(1), 16), addr))
        for lo, hi in sorted(set(loops), key=lambda x: x[0] - x[1]):
            body = [t for a, t in ins if lo <= a <= hi]
            if len(body) < minlen:
                continue
            h = collections.Counter(opcode(t) for t in body)
            print("loop 0x%x..0x%x: %d instructions" % (lo, hi, len(body)))
            print("   " + ", ".join("%s %d" % kv for kv in h.most_common(40)))


if __name__ == "__main__":
    main()
