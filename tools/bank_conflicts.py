import itertools
def patterns(log2n, lgL, nthreads_line):
    """yield (name, list of chunk indices per lane for each warp instruction) for one unit"""
    n = 1 << log2n; n8 = n >> 3; L = 1 << lgL
    lead = log2n % 3
    upt = n8 * L
    pats = []
    def lanes(warp):
        for lane in range(32):
            tu = warp*32 + lane
            yield tu & (L-1), tu >> lgL     # l, lt
    nwarps = max(1, upt // 32)
    for warp in range(nwarps):
        # linear read
        for m in range(8):
            pats.append(('rd', [((lt + m*n8) << lgL) | l for l, lt in lanes(warp)]))
        # partner read
        for m in range(8):
            idx = []
            for l, lt in lanes(warp):
                plt = (n8 - lt) & (n8 - 1); pm = ((8-m)&7) if lt == 0 else (7-m)
                idx.append(((plt + pm*n8) << lgL) | l)
            pats.append(('partner', idx))
        # lead stage writes
        if lead == 2:
            for q in range(2):
                for t in range(4):
                    pats.append(('w_lead4', [((((lt + q*n8) << 2) + t) << lgL) | l for l, lt in lanes(warp)]))
        elif lead == 1:
            for q in range(4):
                for t in range(2):
                    pats.append(('w_lead2', [((((lt + q*n8) << 1) + t) << lgL) | l for l, lt in lanes(warp)]))
        lgNs = lead
        while lgNs < log2n:
            for t in range(8):
                idx = []
                for l, lt in lanes(warp):
                    jlo = lt & ((1 << lgNs) - 1)
                    base = ((lt >> lgNs) << (lgNs + 3)) | jlo
                    idx.append(((base + (t << lgNs)) << lgL) | l)
                pats.append(('w_Ns%d' % (1 << lgNs), idx))
            lgNs += 3
    return pats

def wavefronts(idx, swz):
    tot = 0
    for qw in range(4):
        chunks = set(swz(c) for c in idx[qw*8:(qw+1)*8])
        cnt = {}
        for c in chunks: cnt[c & 7] = cnt.get(c & 7, 0) + 1
        tot += max(cnt.values())
    return tot

def evaluate(log2n, lgL, swz):
    res = {}
    for name, idx in patterns(log2n, lgL, None):
        w = wavefronts(idx, swz)
        a = res.setdefault(name, [0, 0]); a[0] += w; a[1] += 4
    return res

cur = lambda c: c ^ ((c >> 3) & 7)
for (lg, lgL) in [(11, 0), (11, 1), (11, 2), (8, 2), (8, 0)]:
    r = evaluate(lg, lgL, cur)
    tot = sum(v[0] for v in r.values()); ideal = sum(v[1] for v in r.values())
    print('log2n', lg, 'lgL', lgL, 'current swizzle: wavefronts', tot, 'ideal', ideal, {k: round(v[0]/v[1], 2) for k, v in r.items()})
# search: low3 ^= ((c >> s1) & 7) ^ ((c >> s2) & 7) combos
best = {}
for (lg, lgL) in [(11, 0), (11, 1), (11, 2), (8, 2), (8,0)]:
    cands = []
    for s1 in range(1, 9):
        for s2 in [None] + list(range(s1+1, 10)):
            for s3 in [None] + list(range(7, 11)):
                def f(c, s1=s1, s2=s2, s3=s3):
                    x = (c >> s1) & 7
                    if s2 is not None: x ^= (c >> s2) & 7
                    if s3 is not None: x ^= (c >> s3) & 7
                    return c ^ x
                # must be a bijection within aligned 8-blocks? x depends on bits >= s1 >= 1: if s1 < 3, x depends on low bits -> not nec. bijective
                if s1 < 3: 
                    # check bijection on 0..63
                    if len(set(f(c) for c in range(4096))) != 4096: continue
                r = evaluate(lg, lgL, f)
                tot = sum(v[0] for v in r.values())
                cands.append((tot, s1, s2, s3))
    cands.sort()
    print((lg, lgL), 'best', cands[:4])
