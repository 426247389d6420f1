"""Pretty-print the MMA-warp stamps of tools/tc_timeline.py output per item (profiling aid)."""
import sys
L = open(sys.argv[1]).read().split("\n")
def row(name, sec):
    on = False
    for i, l in enumerate(L):
        if l.startswith("====="): on = sec in l
        if on and l.startswith("--- " + name): return list(map(int, L[i + 1].split()))
    return []
for sec in ("COLUMN", "ROW"):
    m = row("mma", sec)
    if not m: continue
    i, k = 1, 0
    print(sec, "first S issued at", m[0])
    while i + 1 < len(m):
        pre, got = m[i], m[i + 1]; i += 2
        ch, s = [], -1
        for n in range(8):
            if n == 2 and i < len(m): s = m[i]; i += 1
            if i + 1 < len(m): ch.append((m[i], m[i + 1])); i += 2
        print(f" item{k}: waitP {pre}->{got} S@{s} chunks", " ".join(f"{a}-{b}" for a, b in ch)); k += 1
    sm = row("softmax", sec)
    print(" softmax (top, S ready, pass1 done, P written):", [tuple(sm[j:j + 4]) for j in range(0, len(sm) - 3, 4)])
    c = row("converter", sec)
    print(" converter (tile ready, done):", [tuple(c[j:j + 2]) for j in range(0, min(len(c), 60) - 1, 2)])
    p = row("producer", sec)
    print(" producer issue:", p[:60])
