#!/usr/bin/env python
"""(CPU, no GPU needed) Model check of the barrier protocol of attn_window_persist_kernel (tc_attention_persist.cuh).

Every mbarrier is modelled as (completed phase count); `wait(bar, parity)` passes iff the barrier's current phase parity
differs from `parity` (PTX try_wait.parity semantics: the phase with that parity has completed).  Each actor is a generator
that yields at every wait; a random scheduler interleaves them.  The check: whenever a wait passes, the event the code means
to wait for (a specific unit's fill / MMA / softmax) has really happened -- i.e. no parity aliasing -- and nobody deadlocks.
"""
import random
import sys


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0
        self.log = []            # payload of each completed phase

    def arrive(self, payload):
        self.pending -= 1
        if self.pending == 0:
            self.pending = self.count
            self.phase += 1
            self.log.append(payload)

    def passes(self, parity):
        return (self.phase & 1) != parity


def run(n_local, seed):
    rng = random.Random(seed)
    full = [Bar(1), Bar(1)]
    s_ready = [Bar(1), Bar(1)]
    p_ready = [Bar(1), Bar(1)]          # 128 threads act in lockstep here: modelled as one arrival
    o_ready = [Bar(1), Bar(1)]
    o_cons = [Bar(1), Bar(1)]
    inflight = []                        # pending async completions: (bar, payload)
    buf_owner = [None, None]             # which unit's data a shared-memory buffer holds
    tmem_owner = [None, None]

    def expect(bar, parity, want_payload):
        while not bar.passes(parity):
            yield
        assert bar.log[-1] == want_payload or want_payload in bar.log[-2:], (bar.log, want_payload)
        assert bar.log[-1] == want_payload, f"parity aliasing: waited for {want_payload}, barrier last completed {bar.log[-1]}"

    def mma():
        def load(i):
            b = i & 1
            assert buf_owner[b] is None or buf_owner[b][1] == "drained", f"refill of live buffer {b}: {buf_owner[b]}"
            buf_owner[b] = (i, "loading")
            inflight.append((full[b], ("fill", i), lambda i=i, b=b: buf_owner.__setitem__(b, (i, "full"))))

        def issue_s(i):
            b = i & 1
            assert buf_owner[b] == (i, "full"), (i, buf_owner)
            assert tmem_owner[b] is None or tmem_owner[b][1] == "consumed", f"S({i}) over live TMEM {tmem_owner[b]}"
            tmem_owner[b] = (i, "S pending")
            inflight.append((s_ready[b], ("S", i), lambda i=i, b=b: tmem_owner.__setitem__(b, (i, "S"))))

        if n_local > 0:
            load(0)
        if n_local > 1:
            load(1)
        if n_local > 0:
            yield from expect(full[0], 0, ("fill", 0))
            issue_s(0)
        for i in range(n_local):
            b, par = i & 1, (i >> 1) & 1
            if i + 1 < n_local:
                b1 = b ^ 1
                yield from expect(full[b1], ((i + 1) >> 1) & 1, ("fill", i + 1))
                if i >= 1:
                    yield from expect(o_cons[b1], ((i - 1) >> 1) & 1, ("Oread", i - 1))
                issue_s(i + 1)
            yield from expect(p_ready[b], par, ("P", i))
            assert tmem_owner[b] == (i, "S read"), tmem_owner
            tmem_owner[b] = (i, "O pending")
            inflight.append((o_ready[b], ("O", i), lambda i=i, b=b: (tmem_owner.__setitem__(b, (i, "O")), buf_owner.__setitem__(b, (i, "drained")))))
            if i + 2 < n_local:
                yield from expect(o_ready[b], par, ("O", i))
                load(i + 2)

    def softmax():
        for i in range(n_local):
            b, par = i & 1, (i >> 1) & 1
            yield from expect(s_ready[b], par, ("S", i))
            assert tmem_owner[b] == (i, "S"), tmem_owner
            assert buf_owner[b] == (i, "full")
            tmem_owner[b] = (i, "S read")
            p_ready[b].arrive(("P", i))
            yield from expect(o_ready[b], par, ("O", i))
            assert tmem_owner[b] == (i, "O"), tmem_owner
            tmem_owner[b] = (i, "consumed")
            o_cons[b].arrive(("Oread", i))
            yield                              # store to global memory

    actors = {"mma": mma(), "softmax": softmax()}
    idle = 0
    while actors:
        # async completions (TMA fills, MMA commits) land in issue order per kind, at random times
        if inflight and rng.random() < 0.5:
            # tcgen05 operations complete in issue order; TMA fills complete in any order relative to them and to each other
            kinds = [0 if it[1][0] == "fill" else 1 for it in inflight]
            first_mma = kinds.index(1) if 1 in kinds else None
            choices = [k for k, kind in enumerate(kinds) if kind == 0] + ([first_mma] if first_mma is not None else [])
            bar, payload, effect = inflight.pop(rng.choice(choices))
            effect()
            bar.arrive(payload)
        name = rng.choice(list(actors))
        before = (tuple(b.phase for b in full + s_ready + p_ready + o_ready + o_cons), len(inflight))
        try:
            next(actors[name])
        except StopIteration:
            del actors[name]
        after = (tuple(b.phase for b in full + s_ready + p_ready + o_ready + o_cons), len(inflight))
        idle = idle + 1 if before == after and not inflight else 0
        assert idle < 10000, f"deadlock with n_local={n_local} seed={seed}: {[(b.phase, b.log[-1:]) for b in full + s_ready + p_ready + o_ready + o_cons]}"
    return True


if __name__ == "__main__":
    n = 0
    for n_local in range(0, 12):
        for seed in range(300):
            run(n_local, seed)
            n += 1
    print(f"attention persist protocol: {n} randomised schedules, no parity aliasing, no deadlock, no live-buffer overwrite")
