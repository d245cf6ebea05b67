#!/usr/bin/env python
"""(CPU) model check of the mbarrier protocols of the persistent tcgen05 kernels:

    gemm(...)   gemm_tc_persist       (k-diffusion_b200/csrc/tc_gemm_persist.cuh): TMA producer, 1-2 MMA issuers with the turn token, NG epilogue groups
    ffn(...)    ffn_fused_kernel      (k-diffusion_b200/csrc/tc_ffn_fused.cuh):    TMA producer, M1 issuer, M2 issuer, three epilogue groups
    attn(...)   attn_pipe_kernel      (k-diffusion_b200/csrc/tc_attention_pipe.cuh): TMA producer, MMA issuer, two softmax groups (and the rejected two-issuer variant)

Agents run the kernels' control flow under a random scheduler.  TMA loads complete asynchronously in any order; the MMAs of ONE
thread complete in its issue order (different threads interleave arbitrarily); a tcgen05.commit arrives when every MMA its thread
issued before it has completed.  mbarrier waits have try_wait.parity semantics.  Checked at every step:
  * no deadlock, every agent terminates;
  * no parity aliasing: the hardware's parity answer equals "phase k has completed" at every wait (a waiter two phases behind, or
    one that the barrier has lapped, fails here -- the first two-issuer GEMM did exactly that on the GPU);
  * no live buffer is overwritten and every consumer sees the data it expects (content tags on ring stages, accumulators, X tiles).
Each model returns True or raises AssertionError.  Used by tests/test_host_logic.py.
"""
import random


class Bar:
    def __init__(self, count=1):
        self.count, self.pending, self.done = count, count, 0          # done = completed phases

    def arrive(self, n=1):
        self.pending -= n
        assert self.pending >= 0, "more arrivals than the barrier expects"
        if self.pending == 0:
            self.pending, self.done = self.count, self.done + 1
            return True
        return False

    def ready(self, k):
        """try_wait.parity for phase k (0-based; k < 0 = the "fresh barrier" wait that must pass immediately)."""
        truth = self.done >= k + 1
        hardware = (self.done & 1) != (k & 1)
        assert hardware == truth, f"parity aliasing: waiting for phase {k}, {self.done} phases completed, the parity test says {hardware}"
        return truth


class Engine:
    def __init__(self, rng):
        self.rng, self.mma, self.tma = rng, {}, []

    def issue(self, thread, fn):                 # an MMA (or a group of MMAs) whose effect `fn` happens at completion
        self.mma.setdefault(thread, []).append(("mma", fn))

    def commit(self, thread, bar, also=None):
        self.mma.setdefault(thread, []).append(("commit", bar, also))

    def load(self, fn):
        self.tma.append(fn)

    def in_flight(self, thread=None):
        return any(q for t, q in self.mma.items() if thread is None or t == thread)

    def pending(self):
        return bool(self.tma) or any(self.mma.values())

    def step(self):
        kinds = [t for t, q in self.mma.items() if q] + (["@tma"] if self.tma else [])
        pick = self.rng.choice(kinds)
        if pick == "@tma":
            self.tma.pop(self.rng.randrange(len(self.tma)))()
            return
        op = self.mma[pick].pop(0)
        if op[0] == "mma":
            op[1]()
        else:
            completed = op[1].arrive()
            if op[2] is not None:
                op[2](completed)


def run_agents(eng, agents, what, max_steps=400000):
    waiting = {k: None for k in agents}
    for _ in range(max_steps):
        if not agents and not eng.pending():
            return True
        ready = [n for n in agents if waiting[n] is None or waiting[n]()]
        if eng.pending():
            ready.append("@async")
        assert ready, f"deadlock: {sorted(agents)} all blocked ({what})"
        pick = eng.rng.choice(ready)
        if pick == "@async":
            eng.step()
            continue
        try:
            waiting[pick] = next(agents[pick])
        except StopIteration:
            del agents[pick], waiting[pick]
    raise AssertionError(f"did not terminate ({what})")


# ------------------------------------------------------------------------------------------------------------------------------
# persistent GEMM
# ------------------------------------------------------------------------------------------------------------------------------
def gemm(n_local, nb, nkb, stages, ng, issuers, seed, b_res=True, token=True):
    """n_local tiles of this CTA; weight-resident mode: nb n-blocks share one A tile (nkb ring stages); streaming: nb must be 1."""
    assert b_res or nb == 1
    rng = random.Random(seed)
    eng = Engine(rng)
    what = f"gemm n_local={n_local} nb={nb} nkb={nkb} stages={stages} ng={ng} issuers={issuers} b_res={b_res} token={token} seed={seed}"
    full = [Bar() for _ in range(stages)]
    empty = [Bar(2 if (b_res and nb >= 2 and issuers == 2) else 1) for _ in range(stages)]
    tmem_full, tmem_empty = [Bar() for _ in range(ng)], [Bar() for _ in range(ng)]
    turn = [Bar(), Bar()]
    stage = [None] * stages                      # (A tile, kb) | ("consumed", ...)
    acc = [None] * ng                            # tile whose MMAs own the accumulator
    acc_done = [None] * ng                       # tile whose MMAs have all completed
    n_atiles = (n_local + nb - 1) // nb

    def producer():
        for a in range(n_atiles):
            for kb in range(nkb):
                pos = a * nkb + kb
                s, k = pos % stages, pos // stages
                yield lambda s=s, k=k: empty[s].ready(k - 1)
                assert stage[s] is None or stage[s][0] == "consumed", f"stage {s} refilled before it was consumed: {stage[s]}"

                def land(s=s, a=a, kb=kb):
                    stage[s] = (a, kb)
                    full[s].arrive()
                eng.load(land)

    def issuer(i_d):
        n = 0
        for it in range(i_d, n_local, issuers):
            a, j = (it // nb, it % nb) if b_res else (it, 0)
            ac, use = it % ng, it // ng
            if issuers == 2 and token:
                yield lambda n=n: turn[i_d].ready(n - 1 if i_d == 0 else n)
            yield lambda ac=ac, use=use: tmem_empty[ac].ready(use - 1)
            first, last = (j < issuers, j + issuers >= nb) if b_res else (True, True)
            for kb in range(nkb):
                pos = a * nkb + kb
                s, k = pos % stages, pos // stages
                if first:
                    yield lambda s=s, k=k: full[s].ready(k)

                def mma(s=s, a=a, kb=kb, it=it, ac=ac):
                    assert stage[s] == (a, kb), f"tile {it}: MMA read stage {s} = {stage[s]}, wanted {(a, kb)}"
                    if kb == 0:
                        assert acc[ac] is None, f"tile {it} started on accumulator {ac} still owned by tile {acc[ac]}"
                        acc[ac] = it
                    assert acc[ac] == it
                eng.issue(i_d, mma)
                if last:
                    def freed(completed, s=s):
                        if completed:
                            stage[s] = ("consumed", stage[s])
                    eng.commit(i_d, empty[s], freed)
                if issuers == 2 and token and kb == nkb - 1:
                    turn[i_d ^ 1].arrive()       # (the kernel signals after its last WAIT; modelled at the last k-block)

            def done(completed, it=it, ac=ac):
                acc_done[ac] = it
            eng.commit(i_d, tmem_full[ac], done)
            n += 1

    def epilogue(g):
        use = 0
        for it in range(g, n_local, ng):
            yield lambda use=use: tmem_full[g].ready(use)
            assert acc_done[g] == it and acc[g] == it, f"group {g} read accumulator of tile {acc[g]} / {acc_done[g]}, wanted {it}"
            yield None                            # tcgen05.ld
            acc[g] = None
            tmem_empty[g].arrive()
            yield None                            # arithmetic, stores
            use += 1

    agents = {"producer": producer(), **{f"issuer{i}": issuer(i) for i in range(issuers)}, **{f"epi{g}": epilogue(g) for g in range(ng)}}
    return run_agents(eng, agents, what)


# ------------------------------------------------------------------------------------------------------------------------------
# fused feed-forward
# ------------------------------------------------------------------------------------------------------------------------------
def ffn(n_local, nc, seed, wu_stages=3, wd_stages=3, ng=3):
    rng = random.Random(seed)
    eng = Engine(rng)
    what = f"ffn n_local={n_local} nc={nc} seed={seed}"
    x_full, x_empty = [Bar(), Bar()], [Bar(), Bar()]
    wu_full, wu_empty = [Bar() for _ in range(wu_stages)], [Bar() for _ in range(wu_stages)]
    wd_full, wd_empty = [Bar() for _ in range(wd_stages)], [Bar() for _ in range(wd_stages)]
    acc1_full, h_ready, acc1_free = [Bar() for _ in range(ng)], [Bar() for _ in range(ng)], [Bar() for _ in range(ng)]
    acc2_full, acc2_free = [Bar() for _ in range(ng)], Bar()
    xbuf, wu, wd = [None, None], [None] * wu_stages, [None] * wd_stages
    acc1 = [None] * ng                           # ("acc", q) after M1, ("H", q) after the group, None when M2 has consumed it
    acc2 = [None, 0]                             # [tile, chunks accumulated]
    acc2_read = [-1]                             # last tile whose acc2 the final epilogue has taken

    def producer():
        def load_x(i):
            buf = i & 1
            yield lambda: x_empty[buf].ready((i >> 1) - 1)
            assert xbuf[buf] is None or xbuf[buf][0] == "stored", f"X buffer {buf} refilled while it holds {xbuf[buf]}"

            def land():
                xbuf[buf] = ("x", i)
                x_full[buf].arrive()
            eng.load(land)
        if n_local:
            yield from load_x(0)
        q = 0
        for i in range(n_local):
            for c in range(nc):
                for ring, fullb, emptyb, n_st in ((wu, wu_full, wu_empty, wu_stages), (wd, wd_full, wd_empty, wd_stages)):
                    s, k = q % n_st, q // n_st
                    yield lambda emptyb=emptyb, s=s, k=k: emptyb[s].ready(k - 1)
                    assert ring[s] is None or ring[s][0] == "consumed", f"weight stage {s} refilled before it was consumed: {ring[s]}"

                    def land(ring=ring, fullb=fullb, s=s, q=q):
                        ring[s] = ("w", q)
                        fullb[s].arrive()
                    eng.load(land)
                q += 1
            if i + 1 < n_local:
                yield from load_x(i + 1)

    def m1():
        q = 0
        for i in range(n_local):
            yield lambda i=i: x_full[i & 1].ready(i >> 1)
            for c in range(nc):
                b, u, s, k = q % ng, q // ng, q % wu_stages, q // wu_stages
                yield lambda b=b, u=u: acc1_free[b].ready(u - 1)
                yield lambda s=s, k=k: wu_full[s].ready(k)

                def mma(i=i, q=q, b=b, s=s):
                    assert xbuf[i & 1] == ("x", i), f"M1 of tile {i} read X buffer {xbuf[i & 1]}"
                    assert wu[s] == ("w", q), f"M1 chunk {q} read Wup stage {wu[s]}"
                    assert acc1[b] is None, f"M1 chunk {q} overwrote accumulator buffer {b} = {acc1[b]}"
                    acc1[b] = ("acc", q)
                eng.issue("m1", mma)
                eng.commit("m1", acc1_full[b])

                def freed(completed, s=s):
                    wu[s] = ("consumed", wu[s])
                eng.commit("m1", wu_empty[s], freed)
                q += 1

    def m2():
        q = 0
        for i in range(n_local):
            yield lambda i=i: acc2_free.ready(i - 1)
            for c in range(nc):
                b, u, s, k = q % ng, q // ng, q % wd_stages, q // wd_stages
                yield lambda s=s, k=k: wd_full[s].ready(k)
                yield lambda b=b, u=u: h_ready[b].ready(u)

                def mma(i=i, c=c, q=q, b=b, s=s):
                    assert acc1[b] == ("H", q), f"M2 chunk {q} read {acc1[b]} from buffer {b}"
                    assert wd[s] == ("w", q), f"M2 chunk {q} read Wdown stage {wd[s]}"
                    if c == 0:
                        assert acc2_read[0] == i - 1, f"acc2 restarted for tile {i} before tile {i - 1} was read"
                        acc2[0], acc2[1] = i, 1
                    else:
                        assert acc2 == [i, c], f"acc2 accumulation out of order: {acc2}"
                        acc2[1] = c + 1
                    acc1[b] = None
                eng.issue("m2", mma)
                eng.commit("m2", acc1_free[b])

                def freed(completed, s=s):
                    wd[s] = ("consumed", wd[s])
                eng.commit("m2", wd_empty[s], freed)
                q += 1
            eng.commit("m2", acc2_full[i % ng])

    def group(g):
        use = 0
        i, c = divmod(g, nc) if g >= nc else (0, g)
        while i < n_local:
            q = i * nc + c
            yield lambda use=use: acc1_full[g].ready(use)
            assert acc1[g] == ("acc", q), f"group {g} read {acc1[g]}, wanted chunk {q}"
            use += 1
            yield None                            # tcgen05.ld + GEGLU
            acc1[g] = ("H", q)
            h_ready[g].arrive()
            if c + ng >= nc and i % ng == g:      # final epilogue of tile i
                buf = i & 1
                yield lambda i=i: acc2_full[g].ready(i // ng)
                assert acc2 == [i, nc], f"final epilogue of tile {i} read acc2 = {acc2}"
                assert x_full[buf].ready(i >> 1)
                yield None
                acc2_read[0] = i
                acc2_free.arrive()
                assert xbuf[buf] == ("x", i) and not eng.in_flight("m1") or xbuf[buf] == ("x", i), "residual read from the wrong X tile"
                yield None                        # residual add in place, TMA store, drain
                xbuf[buf] = ("stored", i)
                x_empty[buf].arrive()
            c += ng
            while c >= nc:
                c -= nc
                i += 1

    agents = {"producer": producer(), "m1": m1(), "m2": m2(), **{f"group{g}": group(g) for g in range(ng)}}
    return run_agents(eng, agents, what)


# ------------------------------------------------------------------------------------------------------------------------------
# pipelined attention
# ------------------------------------------------------------------------------------------------------------------------------
def attn(n_local, nb, shared_kv, seed, stages=5, issuers=1):
    """issuers = 1: the shipped kernel (one thread issues both tiles' MMAs, visiting every K/V stage in order).
    issuers = 2: one issuing thread per tile of the pair, tried in round 2.  With separate K/V per tile (neighbourhood / window modes)
    issuer t then waits only on every second ring position, and with an odd number of stages the previous phase of its stage belongs to
    the OTHER tile: nothing orders that load's completion before the wait (TMA loads may complete out of order), so the parity test
    can pass one phase early and the S MMA would read a stale stage.  The model flags it on the 3-stage ring (on the 5-stage ring a load would have to complete three positions out of order); the GPU tests never hit it, the kernel was
    reverted to one issuer anyway (it was not faster: 0.325 vs 0.315 ms of attention per evaluation)."""
    rng = random.Random(seed)
    eng = Engine(rng)
    what = f"attn n_local={n_local} nb={nb} shared={shared_kv} stages={stages} issuers={issuers} seed={seed}"
    q_full, q_empty = [Bar(), Bar()], [Bar(issuers), Bar(issuers)]
    kv_full, kv_empty = [Bar() for _ in range(stages)], [Bar(2 if (shared_kv and issuers == 2) else 1) for _ in range(stages)]
    s_ready, p_ready, pv_done = [Bar(), Bar()], [Bar(), Bar()], [Bar(), Bar()]
    stage, qbuf = [None] * stages, [None, None]
    S, P, O, o_taken = [None, None], [None, None], [None, None], [-1, -1]

    def producer():
        use = 0
        for u in range(n_local):
            qb = u & 1
            yield lambda: q_empty[qb].ready((u >> 1) - 1)

            def land_q(qb=qb, u=u):
                qbuf[qb] = u
                q_full[qb].arrive()
            eng.load(land_q)
            for j in range(nb):
                for t in ([-1] if shared_kv else [0, 1]):
                    s, k = use % stages, use // stages
                    yield lambda s=s, k=k: kv_empty[s].ready(k - 1)
                    assert stage[s] is None or stage[s][0] == "consumed", f"K/V stage {s} refilled before it was consumed: {stage[s]}"

                    def land(s=s, tag=(u, t, j)):
                        stage[s] = tag
                        kv_full[s].arrive()
                    eng.load(land)
                    use += 1

    def pos_of(u, j, t):
        return (u * nb + j) if shared_kv else (u * nb + j) * 2 + t

    def ops(name, t):
        def issue_s(u, qb, j):
            pos = pos_of(u, j, t)
            s, tag = pos % stages, (u, -1 if shared_kv else t, j)

            def mma():
                assert qbuf[qb] == u, f"S MMA read Q of pair {qbuf[qb]}, wanted {u}"
                assert stage[s] == tag, f"S MMA read K {stage[s]}, wanted {tag}"
                S[t] = (u, j)
            eng.issue(name, mma)
            eng.commit(name, s_ready[t])

        def issue_pv(u, j, release):
            pos = pos_of(u, j, t)
            s, tag = pos % stages, (u, -1 if shared_kv else t, j)

            def pv():
                assert P[t] == (u, j), f"P V MMA read P {P[t]}, wanted {(u, j)}"
                assert stage[s] == tag, f"P V MMA read V {stage[s]}, wanted {tag}"
                if j == 0:
                    assert o_taken[t] >= u - 1, "O_t restarted before the previous pair's result was taken"
                    O[t] = (u, 1)
                else:
                    assert O[t] == (u, j)
                    O[t] = (u, j + 1)
            eng.issue(name, pv)
            eng.commit(name, pv_done[t])
            if release:
                def freed(completed):
                    if completed:
                        stage[s] = ("consumed", stage[s])
                eng.commit(name, kv_empty[s], freed)
        return issue_s, issue_pv

    def wait_kv(u, j, t):
        pos = pos_of(u, j, t)
        return lambda: kv_full[pos % stages].ready(pos // stages)

    def issuer_single():
        fn = [ops("iss", 0), ops("iss", 1)]
        n_p = [0, 0]
        for u in range(n_local):
            qb = u & 1
            yield lambda qb=qb, u=u: q_full[qb].ready(u >> 1)
            for t in (0, 1):
                if t == 0 or not shared_kv:
                    yield wait_kv(u, 0, t)
                fn[t][0](u, qb, 0)
            for j in range(nb):
                more = j + 1 < nb
                for t in (0, 1):
                    yield lambda t=t, k=n_p[t]: p_ready[t].ready(k)
                    n_p[t] += 1
                    fn[t][1](u, j, release=not shared_kv or t == 1)
                    if more:
                        if t == 0 or not shared_kv:
                            yield wait_kv(u, j + 1, t)
                        fn[t][0](u, qb, j + 1)
            eng.commit("iss", q_empty[qb])

    def issuer_per_tile(t):
        name = f"iss{t}"
        issue_s, issue_pv = ops(name, t)
        n_p = 0
        for u in range(n_local):
            qb = u & 1
            yield lambda qb=qb, u=u: q_full[qb].ready(u >> 1)
            yield wait_kv(u, 0, t)
            issue_s(u, qb, 0)
            for j in range(nb):
                yield lambda k=n_p: p_ready[t].ready(k)
                n_p += 1
                issue_pv(u, j, release=True)
                if j + 1 < nb:
                    yield wait_kv(u, j + 1, t)
                    issue_s(u, qb, j + 1)
            eng.commit(name, q_empty[qb])

    def softmax(t):
        n = 0
        for u in range(n_local):
            for j in range(nb):
                yield lambda n=n: s_ready[t].ready(n)
                assert S[t] == (u, j), f"group {t} read S {S[t]}, wanted {(u, j)}"
                yield None
                P[t] = (u, j)
                p_ready[t].arrive()
                n += 1
            yield lambda n=n: pv_done[t].ready(n - 1)
            assert O[t] == (u, nb), f"group {t} read O {O[t]}, wanted {(u, nb)}"
            o_taken[t] = u
            yield None

    agents = {"producer": producer(), "softmax0": softmax(0), "softmax1": softmax(1)}
    if issuers == 1:
        agents["iss"] = issuer_single()
    else:
        agents.update(iss0=issuer_per_tile(0), iss1=issuer_per_tile(1))
    return run_agents(eng, agents, what)


if __name__ == "__main__":
    total = 0
    for seed in range(40):
        for nb, nkb, stages, ng in ((3, 2, 4, 2), (3, 2, 4, 3), (1, 2, 6, 2), (1, 4, 6, 3), (2, 2, 6, 2), (1, 6, 6, 2)):
            for issuers in (1, 2):
                for n_local in (0, 1, 2, 3, 7, 12):
                    gemm(n_local * nb if n_local else 0, nb, nkb, stages, ng, issuers, seed)
                    total += 1
        for nkb, stages, ng in ((8, 4, 2), (32, 4, 3), (12, 4, 2)):
            for issuers in (1, 2):
                for n_local in (1, 2, 5):
                    gemm(n_local, 1, nkb, stages, ng, issuers, seed, b_res=False)
                    total += 1
        for nc in (3, 4, 6, 8):
            for n_local in (0, 1, 2, 3, 7):
                ffn(n_local, nc, seed)
                total += 1
        for shared in (True, False):
            for nb in (1, 2, 3, 8):
                for n_local in (0, 1, 2, 5):
                    attn(n_local, nb, shared, seed)
                    attn(n_local, nb, shared, seed, stages=3)
                    total += 2
    print(f"{total} random schedules: no deadlock / parity aliasing / overwrite in the GEMM, fused feed-forward and attention protocols")
