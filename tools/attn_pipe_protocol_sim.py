#!/usr/bin/env python
"""(CPU) model check of the mbarrier protocol of attn_pipe_kernel (k-diffusion_b200/csrc/tc_attention_pipe.cuh).

Four agents -- TMA producer, tcgen05.mma issuer, softmax group 0, softmax group 1 -- run the kernel's control flow under a random
scheduler.  TMA loads and MMAs complete asynchronously (MMAs in issue order, as the tensor pipe does; a commit arrives when every
MMA issued before it has completed).  mbarrier waits use try_wait.parity semantics.  Checked on every step:
  * no deadlock, every agent terminates;
  * no parity aliasing: a wait for phase k of a barrier only ever sees k or k + 1 completed phases (never passes early, never misses);
  * no live buffer is overwritten: K/V stage refilled while an MMA that reads it is in flight or before it was consumed, Q buffer
    refilled under a running S MMA, S_t rewritten while its softmax group still reads it, P_t rewritten under a running P V MMA,
    O_t restarted before the group has taken the previous pair's result;
  * every consumer sees the data it expects (S, P, K/V, Q content tags).
run(n_local, nb, shared_kv, seed) -> True or raises AssertionError.  Used by tests/test_host_logic.py.
"""
import random

STAGES = 3


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.done = count, count, 0          # done = completed phases

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.pending, self.done = self.count, self.done + 1

    def ready(self, k):
        """try_wait.parity for phase k (0-based).  Aliasing = the hardware test would answer for another phase."""
        assert self.done in (k, k + 1) or self.done < k, f"barrier ran ahead: waiting for phase {k}, {self.done} completed"
        passes = (self.done & 1) != (k & 1)                    # parity of the current phase differs from the waited one
        if self.done < k:
            assert not passes, f"parity aliasing: wait for phase {k} passes with {self.done} completed"
        return passes if self.done <= k + 1 else False


class Sim:
    def __init__(self, n_local, nb, shared, rng):
        self.n, self.nb, self.shared, self.rng = n_local, nb, shared, rng
        B = lambda c: Bar(c)
        self.q_full, self.q_empty = [B(1), B(1)], [B(1), B(1)]
        self.kv_full, self.kv_empty = [B(1) for _ in range(STAGES)], [B(1) for _ in range(STAGES)]
        self.s_ready, self.p_ready = [B(1), B(1)], [B(1), B(1)]       # p_ready: one arrival models the 128 threads of a group
        self.pv_done = [B(1), B(1)]
        self.stage = [None] * STAGES          # content tag (pair, t, j) ; t = -1 for shared
        self.stage_loading = [False] * STAGES
        self.qbuf = [None, None]
        self.qbuf_loading = [False, False]
        self.S = [None, None]                 # tag (pair, j) once the S MMA has COMPLETED
        self.P = [None, None]
        self.O = [None, None]                 # (pair, blocks accumulated)
        self.s_reading = [False, False]
        self.p_writing = [False, False]
        self.o_taken = [-1, -1]               # last pair whose O the group has read
        self.mma_queue = []                   # in-order: dict(kind, ...), executes (reads/writes) at completion
        self.tma_queue = []

    # ---------------------------------------------------------------- async engines
    def mma_reads(self, pred):
        return any(pred(m) for m in self.mma_queue)

    def step_mma(self):
        m = self.mma_queue.pop(0)
        if m["kind"] == "S":
            t = m["t"]
            assert self.qbuf[m["qb"]] == m["pair"], f"S MMA read Q of pair {self.qbuf[m['qb']]}, wanted {m['pair']}"
            assert self.stage[m["stage"]] == m["kv"], f"S MMA read K {self.stage[m['stage']]}, wanted {m['kv']}"
            assert not self.s_reading[t], "S_t rewritten while its softmax group reads it"
            self.S[t] = (m["pair"], m["j"])
        elif m["kind"] == "PV":
            t = m["t"]
            assert self.P[t] == (m["pair"], m["j"]), f"P V MMA read P {self.P[t]}, wanted {(m['pair'], m['j'])}"
            assert not self.p_writing[t], "P_t rewritten under a running P V MMA"
            assert self.stage[m["stage"]] == m["kv"], f"P V MMA read V {self.stage[m['stage']]}, wanted {m['kv']}"
            if m["j"] == 0:
                assert self.o_taken[t] >= m["pair_index"] - 1, "O_t restarted before the previous pair's result was taken"
                self.O[t] = (m["pair"], 1)
            else:
                assert self.O[t] == (m["pair"], m["j"]), f"O_t accumulation out of order: {self.O[t]}"
                self.O[t] = (m["pair"], m["j"] + 1)
        elif m["kind"] == "commit":
            m["bar"].arrive()
            if m.get("frees_stage") is not None:
                self.stage[m["frees_stage"]] = ("consumed", self.stage[m["frees_stage"]])

    def step_tma(self):
        i = self.rng.randrange(len(self.tma_queue))
        m = self.tma_queue.pop(i)
        if m["kind"] == "kv":
            self.stage[m["stage"]] = m["tag"]
            self.stage_loading[m["stage"]] = False
            self.kv_full[m["stage"]].arrive()
        else:
            self.qbuf[m["qb"]] = m["pair"]
            self.qbuf_loading[m["qb"]] = False
            self.q_full[m["qb"]].arrive()

    # ---------------------------------------------------------------- agents (generators: yield a predicate to wait on, or None)
    def producer(self):
        st, use = 0, 0
        for u in range(self.n):
            qb = u & 1
            yield lambda: self.q_empty[qb].ready((u >> 1) - 1) if u >= 2 else True
            assert not self.mma_reads(lambda m: m["kind"] == "S" and m["qb"] == qb), "Q buffer refilled under a running S MMA"
            self.qbuf_loading[qb] = True
            self.tma_queue.append(dict(kind="q", qb=qb, pair=u))
            for j in range(self.nb):
                for t in ([-1] if self.shared else [0, 1]):
                    k = use // STAGES
                    s = st
                    yield lambda s=s, k=k: self.kv_empty[s].ready(k - 1) if k >= 1 else True
                    assert not self.mma_reads(lambda m: m.get("stage") == s and m["kind"] in ("S", "PV")), "K/V stage refilled under a running MMA"
                    assert self.stage[s] is None or self.stage[s][0] == "consumed", f"K/V stage {s} refilled before it was consumed: {self.stage[s]}"
                    self.stage_loading[s] = True
                    self.tma_queue.append(dict(kind="kv", stage=s, tag=(u, t, j)))
                    st = (st + 1) % STAGES
                    use += 1

    def mma(self):
        cursor = [0, 0]          # (slot, uses)
        n_p = [0, 0]

        def take():
            s, k = cursor[0], cursor[1] // STAGES
            cursor[0] = (cursor[0] + 1) % STAGES
            cursor[1] += 1
            return s, k

        def issue_s(t, u, qb, slot, j):
            tag = (u, -1 if self.shared else t, j)
            self.mma_queue.append(dict(kind="S", t=t, pair=u, qb=qb, stage=slot, kv=tag, j=j))
            self.mma_queue.append(dict(kind="commit", bar=self.s_ready[t]))

        for u in range(self.n):
            qb = u & 1
            cur = [take()]
            cur.append(cur[0] if self.shared else take())
            yield lambda: self.q_full[qb].ready(u >> 1)
            for t in (0, 1):
                if t == 0 or not self.shared:
                    s, k = cur[t]
                    yield lambda s=s, k=k: self.kv_full[s].ready(k)
                issue_s(t, u, qb, cur[t][0], 0)
            for j in range(self.nb):
                more = j + 1 < self.nb
                nxt = None
                if more:
                    nxt = [take()]
                    nxt.append(nxt[0] if self.shared else take())
                for t in (0, 1):
                    k = n_p[t]
                    yield lambda t=t, k=k: self.p_ready[t].ready(k)
                    n_p[t] += 1
                    tag = (u, -1 if self.shared else t, j)
                    self.mma_queue.append(dict(kind="PV", t=t, pair=u, pair_index=u, j=j, stage=cur[t][0], kv=tag))
                    self.mma_queue.append(dict(kind="commit", bar=self.pv_done[t]))
                    if not self.shared or t == 1:
                        self.mma_queue.append(dict(kind="commit", bar=self.kv_empty[cur[t][0]], frees_stage=cur[t][0]))
                    if more:
                        if t == 0 or not self.shared:
                            s, k2 = nxt[t]
                            yield lambda s=s, k2=k2: self.kv_full[s].ready(k2)
                        issue_s(t, u, qb, nxt[t][0], j + 1)
                cur = nxt
            self.mma_queue.append(dict(kind="commit", bar=self.q_empty[qb]))

    def softmax(self, t):
        n = 0
        for u in range(self.n):
            for j in range(self.nb):
                yield lambda n=n: self.s_ready[t].ready(n)
                assert self.S[t] == (u, j), f"group {t} read S {self.S[t]}, wanted {(u, j)}"
                self.s_reading[t] = True
                yield None                                                   # chunk 0 arithmetic
                assert not self.mma_reads(lambda m: m["kind"] == "PV" and m["t"] == t), "P_t rewritten under a running P V MMA"
                self.p_writing[t] = True
                yield None
                yield None
                self.P[t] = (u, j)
                self.p_writing[t] = False
                self.s_reading[t] = False
                self.p_ready[t].arrive()
                n += 1
            yield lambda n=n: self.pv_done[t].ready(n - 1)
            assert self.O[t] == (u, self.nb), f"group {t} read O {self.O[t]}, wanted {(u, self.nb)}"
            self.o_taken[t] = u
            yield None                                                       # global stores


def run(n_local, nb, shared_kv, seed, max_steps=200000):
    rng = random.Random(seed)
    sim = Sim(n_local, nb, shared_kv, rng)
    agents = {"producer": sim.producer(), "mma": sim.mma(), "softmax0": sim.softmax(0), "softmax1": sim.softmax(1)}
    waiting = {k: None for k in agents}
    for _ in range(max_steps):
        if not agents and not sim.mma_queue and not sim.tma_queue:
            return True
        choices = []
        for name in agents:
            w = waiting[name]
            if w is None or w():
                choices.append(name)
        if sim.mma_queue:
            choices.append("@mma")
        if sim.tma_queue:
            choices.append("@tma")
        assert choices, f"deadlock: agents {list(agents)} all blocked (n_local={n_local} nb={nb} shared={shared_kv} seed={seed})"
        pick = rng.choice(choices)
        if pick == "@mma":
            sim.step_mma()
        elif pick == "@tma":
            sim.step_tma()
        else:
            try:
                waiting[pick] = next(agents[pick])
            except StopIteration:
                del agents[pick]
                del waiting[pick]
    raise AssertionError("did not terminate")


if __name__ == "__main__":
    total = 0
    for shared in (True, False):
        for nb in (1, 2, 3, 8):
            for n_local in (0, 1, 2, 3, 5, 8):
                for seed in range(60):
                    run(n_local, nb, shared, seed)
                    total += 1
    print(f"attn_pipe protocol: {total} random schedules, no deadlock / aliasing / overwrite")
