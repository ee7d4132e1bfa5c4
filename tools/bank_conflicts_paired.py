"""Shared-memory wavefront count of the paired-final-stage transforms (fft_core.h: fft_line_post /
fft_line_pre) per access pattern, with the swizzle the kernels use.  Ideal = 4 wavefronts per warp-wide
16-byte access (one per quarter warp)."""
import sys

def swz(c, lgL):
    return c ^ ((c >> 3) & 7) if lgL == 0 else c ^ ((c >> 2) & 6)

def lanes(warp, lgL):
    L = 1 << lgL
    for lane in range(32):
        tu = warp * 32 + lane
        yield tu & (L - 1), tu >> lgL

def wavefronts(idx, lgL, raw=False):
    tot = 0
    for qw in range(4):
        chunks = set((c if raw else swz(c, lgL)) for c in idx[qw * 8:(qw + 1) * 8])
        cnt = {}
        for c in chunks:
            cnt[c & 7] = cnt.get(c & 7, 0) + 1
        tot += max(cnt.values())
    return tot

def evaluate(LG, lgL):
    n8, n4 = 1 << (LG - 3), 1 << (LG - 2)
    upt = n8 << lgL
    res = {}
    def add(name, idx, raw=False):
        a = res.setdefault(name, [0, 0]); a[0] += wavefronts(idx, lgL, raw); a[1] += 4
    for warp in range(max(1, upt // 32)):
        ln = list(lanes(warp, lgL))
        jB = lambda lt: (n4 - lt) if lt else n8
        for m in range(8):
            add('rd_std', [((lt + m * n8) << lgL) | l for l, lt in ln])
        for t in range(4):
            add('rd_pairA', [((lt + t * n4) << lgL) | l for l, lt in ln])
            add('rd_pairB', [((jB(lt) + t * n4) << lgL) | l for l, lt in ln])
            add('w_pre4A', [((4 * lt + t) << lgL) | l for l, lt in ln])
            add('w_pre4B', [((4 * jB(lt) + t) << lgL) | l for l, lt in ln])
        for lgNs in list(range(0, LG - 2, 3)) + list(range(2, LG - 3, 3)):
            for t in range(8):
                idx = []
                for l, lt in ln:
                    jlo = lt & ((1 << lgNs) - 1)
                    base = ((lt >> lgNs) << (lgNs + 3)) | jlo
                    idx.append(((base + (t << lgNs)) << lgL) | l)
                add('w_Ns%d' % (1 << lgNs), idx)
    return res

for LG, lgL in [(11, 0), (11, 1), (11, 2), (8, 0), (8, 2), (8, 3), (5, 0)]:
    r = evaluate(LG, lgL)
    print('LG', LG, 'lgL', lgL, {k: round(v[0] / v[1], 2) for k, v in r.items()})
