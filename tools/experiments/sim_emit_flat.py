import random
def clz32(x): return 32 - x.bit_length()
def ffs(x): return (x & -x).bit_length()
def sim(seed):
    rnd = random.Random(seed)
    # runs
    runs = []; n = 0
    for _ in range(rnd.randint(1, 300)):
        ln = rnd.choice([1,1,1,2,3,5,9,40,70, rnd.randint(1,200)])
        runs.append((n, n+ln)); n += ln
    uniq = [rnd.random() < 0.8 for _ in range(n)]
    cand_runs = [r for r in range(len(runs)) if sum(uniq[runs[r][0]:runs[r][1]]) >= 2 and rnd.random() < 0.7]
    nc = len(cand_runs)
    if nc == 0: return True
    # reference
    ref = []
    for r in cand_runs:
        s,e = runs[r]; ref.append([i for i in range(s,e) if uniq[i]])
    offs = [0]
    for l in ref: offs.append(offs[-1]+len(l))
    out = [None]*offs[-1]; first = [None]*nc; hs = [0]*nc
    INF = 0xFFFFFFFF
    for c0 in range(0, nc, 32):
        S = [INF]*32; E=[INF]*32; OFF=[0]*32; CNT=[0]*32; X0=[None]*32
        for l in range(32):
            c = c0+l
            if c < nc: S[l],E[l] = runs[cand_runs[c]]; OFF[l]=offs[c]
        nvalid = min(32, nc-c0)
        i_begin = S[0]; i_end = E[nvalid-1]
        i0 = i_begin
        while i0 < i_end:
            lanes = []
            for lane in range(32):
                i = i0+lane; inn = i < i_end; j=0; ok=False
                if inn:
                    for step in (16,8,4,2,1):
                        if S[j+step] <= i: j += step
                    ok = i < E[j]
                sel = ok and uniq[i]
                lanes.append((i,inn,j,ok,sel))
            segid = [ (l[2] if l[3] else INF) for l in lanes]
            heads = 0; selb = 0
            for lane in range(32):
                if lane == 0 or segid[lane-1] != segid[lane]: heads |= 1<<lane
                if lanes[lane][4]: selb |= 1<<lane
            ks = [0]*32
            for lane in range(32):
                i,inn,j,ok,sel = lanes[lane]
                le = 0xFFFFFFFF >> (31-lane)
                seg_start = 31 - clz32(heads & le)
                before = selb & ((1<<lane)-1) & ~((1<<seg_start)-1) & 0xFFFFFFFF
                if sel:
                    ks[lane] = CNT[j] + bin(before).count('1')
                    if ks[lane] == 0: X0[j] = i
            newcnt = {}
            for lane in range(32):
                i,inn,j,ok,sel = lanes[lane]
                if sel:
                    le = 0xFFFFFFFF >> (31-lane)
                    k = ks[lane]
                    out[OFF[j]+k] = i
                    hs[c0+j] += i*2654435761 + (X0[j] or 0)
                    above = (~le) & 0xFFFFFFFF
                    nh = heads & above
                    seg_mask = ((1 << (ffs(nh)-1)) - 1) if nh else 0xFFFFFFFF
                    if (selb & above & seg_mask) == 0: newcnt[j] = k+1
            for j,v in newcnt.items(): CNT[j] = v
            anyok = any(l[3] for l in lanes)
            nxt = i0+32
            if not anyok:
                m = min([S[l] if S[l] > i0+31 else INF for l in range(32)])
                if m == INF: break
                nxt = max(nxt, m)
            i0 = nxt
        for l in range(nvalid): first[c0+l] = X0[l]
    flat = [x for l in ref for x in l]
    assert out == flat, (seed,)
    assert first == [l[0] for l in ref]
    assert hs == [sum(i*2654435761 + l[0] for i in l) for l in ref]
    return True
for sd in range(3000): sim(sd)
print("ok")
