"""Discrete-event model of the single-launch forward schedulers (CPU only): 148 CTAs, B samples x (W column + H row) lines,
a row line of sample b becomes available when all column lines of b have finished.  Reports the makespan and the mean
'reuse distance' (bytes of other lines touched between a sample's column pass and its row pass) for a few policies."""
import heapq, sys
NCTA, B, W, H = 148, 8, 97, 97
TC, TR = 15.3, 20.0          # kilo-cycles per column / row line when the operands hit L2
FETCH = 0.25                 # a CTA decides its next line this far into the current one
LINE_MB = 0.287 * 1.0        # q,k,v (+partial) bytes a line touches
def simulate(policy, cap=1):
    t_free = [0.0] * NCTA
    done_cols = [0] * B; col_done_t = [None] * B
    next_col = 0; next_row = [0] * B
    ev = [(0.0, c) for c in range(NCTA)]   # (time the CTA fetches its next line, cta)
    heapq.heapify(ev)
    pending = []                           # (finish time, sample) of running column lines
    finish = 0.0; row_start = [None] * B; idle = 0.0
    first_open = 0
    while ev:
        t, c = heapq.heappop(ev)
        # retire finished columns
        pending.sort()
        while pending and pending[0][0] <= t:
            ft, b = pending.pop(0); done_cols[b] += 1
            if done_cols[b] == W: col_done_t[b] = ft
        while first_open < B and next_row[first_open] >= H: first_open += 1
        item = None
        for b in range(first_open, B):
            if done_cols[b] < W: break
            if next_row[b] < H:
                item = ("r", b); next_row[b] += 1
                if row_start[b] is None: row_start[b] = max(t, t_free[c])
                break
        if item is None and next_col < B * W:
            b = next_col // W
            if policy == "dynamic" or b <= first_open + cap:
                item = ("c", b); next_col += 1
        if item is None:
            if first_open >= B and next_col >= B * W: continue
            # nothing available: retry when the next column finishes
            nt = min([ft for ft, _ in pending if ft > t] + [t + 1.0])
            heapq.heappush(ev, (nt, c)); continue
        start = max(t, t_free[c]); idle += max(0.0, t - t_free[c]) if t_free[c] > 0 else 0.0
        dur = TC if item[0] == "c" else TR
        end = start + dur; t_free[c] = end; finish = max(finish, end)
        if item[0] == "c": pending.append((end, item[1]))
        heapq.heappush(ev, (start + dur * (1 - FETCH) if FETCH else end, c))
    lag = [row_start[b] - col_done_t[b] for b in range(B)]
    return finish, idle / NCTA, lag
ideal = B * (W * TC + H * TR) / NCTA
print("ideal (perfect balance): %.1f Kclk" % ideal)
for pol, cap in (("dynamic", 0), ("capped", 0), ("capped", 1), ("capped", 2)):
    f, idle, lag = simulate(pol, cap)
    print("%-8s cap=%d  makespan %.1f Kclk (%.0f%% of ideal)  idle/CTA %.1f  row-start lag per sample %s" % (pol, cap, f, 100 * f / ideal, idle, [round(x, 1) for x in lag]))
