"""Event simulation of the imma kernel ring protocol (full mbarriers with phase parity, per-slot release counters, refill by
the last releaser) for the single-tile and the paired-tile loop: no deadlock, phase mismatch or premature overwrite for
S in {2,3}, 1-7 tiles per strip, 1-3 strips (used to rule the protocol out as the cause of the paired variant hang)."""
import random, itertools
def simulate(T, S, tiles, pair, nwarps=16, seed=0, latency=(3,9)):
    """event simulation of the imma kernel's ring protocol (full mbarriers with phases, release counters, last-releaser refill)"""
    rnd = random.Random(seed)
    full_phase = [0]*S           # number of completed fills per slot
    pending = []                 # (arrival_time, slot)
    released = [0]*S
    issued_tiles = []            # order of tile requests (slot)
    time = 0
    def issue(slot, tile):
        issued_tiles.append((tile, slot))
        pending.append((time + rnd.randint(*latency), slot))
    # prologue
    r = 0
    for t in range(min(S,T)):
        issue(t, r); r += 1
    class W:
        def __init__(s, wid):
            s.wid=wid; s.t=0; s.kt=0; s.slot=0; s.ph=0; s.r=min(S,T); s.state='run'; s.waits=[]; s.done=False
    ws=[W(i) for i in range(nwarps)]
    def step(w):
        nonlocal time
        if w.done: return False
        if w.t >= T: w.done=True; return True
        two = pair and (w.kt + 1 < tiles)
        slots=[w.slot]; phs=[w.ph]; tls=[w.t]
        if two:
            s1=(w.slot+1)%S; slots.append(s1); phs.append(w.ph ^ 1 if w.slot+1==S else w.ph); tls.append(w.t+1)
        # wait: phase parity wait: barrier completed phase count must be > number implied: completed fills of slot must be >= (tile//S)+1
        for sl,tl,ph in zip(slots,tls,phs):
            need = tl//S + 1
            # parity semantic: wait(parity ph) passes if current completed count parity... emulate exact count but also check parity consistency
            if full_phase[sl] < need: return False
            assert (need-1) % 2 == ph, ("phase mismatch", tl, sl, ph, need)
            assert full_phase[sl] == need, ("slot overwritten before consumption?", tl, sl, full_phase[sl], need)
        # compute done; release
        for i,(sl,tl) in enumerate(zip(slots,tls)):
            released[sl]+=1
            last = released[sl] % nwarps == 0
            if last and tl + S < T:
                issue(sl, w.r)     # every warp tracks r identically
                assert w.r == tl + S, ("refill target mismatch", w.r, tl+S)
            w.r += 1
        adv=len(slots)
        w.slot += adv
        if w.slot >= S: w.slot -= S; w.ph ^= 1
        w.t += adv; w.kt += adv
        if w.kt == tiles: w.kt = 0
        return True
    idle=0
    while not all(w.done for w in ws):
        progressed=False
        order=list(range(nwarps)); rnd.shuffle(order)
        for i in order:
            if rnd.random()<0.7: progressed |= step(ws[i])
        # deliver arrivals
        time+=1
        for a in sorted([p for p in pending if p[0]<=time]):
            full_phase[a[1]] += 1; pending.remove(a); progressed=True
        idle = 0 if progressed else idle+1
        if idle>200: return "DEADLOCK", [ (w.t,w.slot) for w in ws], full_phase, pending
    return "ok", issued_tiles
for pair in (0,1):
    for S in (2,3):
        for tiles,strips in ((2,1),(2,2),(2,3),(6,1),(7,1),(3,2),(1,3),(4,2)):
            T=tiles*strips
            if S>T: continue
            for seed in range(20):
                res=simulate(T,S,tiles,pair,seed=seed)
                if res[0]!="ok": print("pair",pair,"S",S,"tiles",tiles,"strips",strips,res); break
print("done")
