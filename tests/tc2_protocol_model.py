"""A model of the synchronisation protocol of repulse_tc2_kernel (topolow_b200/csrc/rowblock_tc2.cuh) - test
infrastructure.  compute-sanitizer is closed on the GPU pool, so the hand-off protocol of the kernel's ten warps (one
copy warp, one MMA warp, eight consumer warps), its asynchronous engines (bulk copies, the in-order tensor pipe) and its
mbarriers is restated here as a discrete-event model and run under random interleavings.  The model follows the kernel's
control flow statement by statement (cursors, phase parities, who arrives where); data are replaced by version tags, and
every read checks that it sees the version it was meant to see:

  * a wait that can never end (deadlock), or a barrier that ran two phases ahead of a waiter (on the hardware: a hang),
  * a shared-memory stage, an S / w buffer, the D2 accumulator, the rows' image (shared and tensor memory) or an entry of
    the item ring overwritten before its readers were done, or read before it was written.

`run(seed, items, ...)` raises AssertionError on any of them."""
import random

STAGES, BUFS, CONSUMERS = 4, 2, 8


class Barrier:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than its count"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def ready(self, k):
        """try_wait.parity(k & 1) as issued by a waiter that means phase k (k = -1: the fresh-barrier idiom)."""
        assert self.phase <= k + 2 - 1, f"{self.name}: ran ahead of a waiter for phase {k} (now {self.phase} complete)"
        return (self.phase & 1) != (k & 1)


class Model:
    def __init__(self, seed, item_shapes, skip_prob=0.3, first_item_guard=True):
        self.first_item_guard = first_item_guard
        self.rng = random.Random(seed)
        self.shapes = item_shapes                    # item id -> list of stage counts per chunk
        self.n_items = len(item_shapes)
        self.counter = 0                             # dv.counters[2]; other CTAs draw from it too
        self.skip_prob = skip_prob
        B = Barrier
        self.a_full, self.a_free, self.a_tmem, self.item = B("a_full", 1), B("a_free", 1), B("a_tmem", CONSUMERS), B("item", 1)
        self.full = [B(f"full{s}", 1) for s in range(STAGES)]
        self.empty = [B(f"empty{s}", CONSUMERS + 1) for s in range(STAGES)]
        self.tfull = [B(f"tfull{b}", 1) for b in range(BUFS)]
        self.wfull = [B(f"wfull{b}", CONSUMERS) for b in range(BUFS)]
        self.d2full, self.d2empty = B("d2full", 1), B("d2empty", CONSUMERS)
        # data as version tags
        self.ring = [None, None]
        self.stage = [None] * STAGES                 # global stage index whose operands it holds
        self.stage_readers = [set() for _ in range(STAGES)]
        self.a_smem = None                           # item sequence number
        self.a_tmem_parts = [None] * CONSUMERS
        self.sbuf = [("free", -1)] * BUFS            # ("S", g) | ("w", g) | ("free", g)
        self.w_parts = [set() for _ in range(BUFS)]
        self.d2 = ("read", -1)                       # ("acc", chunk) | ("full", chunk) | ("read", chunk)
        self.d2_readers = set()
        self.tma = []                                # pending bulk copies: callables
        self.pipe = []                               # tensor pipe, in order: callables
        self.gemm1_left = {}                         # item seq -> GEMM 1 ops issued and not yet executed
        self.stages_done = 0
        self.speed = {}

    # ---- helpers -------------------------------------------------------------------------------
    def draw_id(self):
        while self.rng.random() < self.skip_prob:    # another CTA took one
            self.counter += 1
        self.counter += 1
        return self.counter - 1

    def chunks_of(self, item):
        return self.shapes[item]

    # ---- roles (generators: yield a predicate to wait for, or None to give way) ----------------
    def copy_warp(self):
        seq = load_g = item_g = 0

        def draw():
            nonlocal seq
            if seq > 0:
                yield lambda k=seq - 1: self.a_tmem.ready(k)
            if seq == 1 and self.first_item_guard:
                yield lambda: self.a_free.ready(0)
            ident = self.draw_id()
            self.ring[seq & 1] = (ident, seq)
            self.item.arrive()
            seq += 1
            return ident
        ident = yield from draw()
        while ident < self.n_items:
            for ci, nst in enumerate(self.chunks_of(ident)):
                for st in range(nst):
                    if ci == 0 and st == 0:
                        if item_g > 0:
                            yield lambda k=item_g - 1: self.a_free.ready(k)
                        item_g += 1
                        assert self.gemm1_left.get(item_g - 2, 0) == 0, "A tile reloaded under a running GEMM 1"

                        def land_a(k=item_g - 1):
                            self.a_smem = k
                            self.a_full.arrive()
                        self.tma.append(land_a)
                    s = load_g % STAGES
                    yield lambda k=load_g // STAGES - 1, s=s: self.empty[s].ready(k)
                    assert self.stage[s] is None or self.stage_readers[s] == set(range(CONSUMERS)) | {"g1", "g2"}, \
                        f"stage {s} refilled before its readers were done: {self.stage_readers[s]}"

                    def land(s=s, g=load_g):
                        self.stage[s] = g
                        self.stage_readers[s] = set()
                        self.full[s].arrive()
                    self.stage[s] = "loading"
                    self.tma.append(land)
                    load_g += 1
                    yield None
            ident = yield from draw()

    def mma_warp(self):
        taken = [None, None]
        g1 = dict(seq=0, item=None, ci=0, st=0, valid=False)
        g2 = dict(seq=0, item=None, ci=0, st=0, valid=False)
        g1_g = g2_g = a_uses = chunk_g = 0

        def open_(cu, ident):
            cu.update(item=ident, ci=0, st=0, valid=ident < self.n_items)

        def advance(cu):
            cu["st"] += 1
            if cu["st"] < self.chunks_of(cu["item"])[cu["ci"]]:
                return
            cu["st"] = 0
            cu["ci"] += 1
            if cu["ci"] < len(self.chunks_of(cu["item"])):
                return
            cu["valid"] = False

        def take_g1():
            yield lambda k=g1["seq"]: self.item.ready(k)
            ident, tag = self.ring[g1["seq"] & 1]
            assert tag == g1["seq"], "MMA warp read a stale / overwritten ring entry"
            g1["seq"] += 1
            open_(g1, ident)
            taken[(g1["seq"] - 1) & 1] = ident

        def gemm1():
            nonlocal g1_g, a_uses
            if g1["ci"] == 0 and g1["st"] == 0:
                yield lambda k=a_uses: self.a_full.ready(k)
                yield lambda k=a_uses: self.a_tmem.ready(k)
                a_uses += 1
            s, b, g, it = g1_g % STAGES, g1_g % BUFS, g1_g, a_uses - 1
            yield lambda k=g1_g // STAGES, s=s: self.full[s].ready(k)
            last_of_item = g1["st"] + 1 == self.chunks_of(g1["item"])[g1["ci"]] and g1["ci"] + 1 == len(self.chunks_of(g1["item"]))
            self.gemm1_left[it] = self.gemm1_left.get(it, 0) + 1

            def op():
                assert self.stage[s] == g, f"GEMM 1 of stage {g} read stage slot {s} holding {self.stage[s]}"
                assert self.a_smem == it, f"GEMM 1 of item {it} read the shared-memory rows of item {self.a_smem}"
                assert self.a_tmem_parts == [it] * CONSUMERS, f"GEMM 1 of item {it} read tensor-memory rows {self.a_tmem_parts}"
                assert self.sbuf[b] == ("free", g - BUFS) or (g < BUFS and self.sbuf[b] == ("free", -1)), \
                    f"GEMM 1 of stage {g} overwrote buffer {b} in state {self.sbuf[b]}"
                self.sbuf[b] = ("S", g)
                self.w_parts[b] = set()
                self.stage_readers[s].add("g1")
                self.gemm1_left[it] -= 1
            self.pipe.append(op)
            self.pipe.append(self.tfull[b].arrive)
            if last_of_item:
                self.pipe.append(self.a_free.arrive)
            g1_g += 1
            advance(g1)
            if not g1["valid"]:
                yield from take_g1()

        def gemm2():
            nonlocal g2_g, chunk_g
            s, b, g = g2_g % STAGES, g2_g % BUFS, g2_g
            first = g2["st"] == 0
            last = g2["st"] + 1 == self.chunks_of(g2["item"])[g2["ci"]]
            yield lambda k=g2_g // BUFS, b=b: self.wfull[b].ready(k)
            if first:
                yield lambda k=chunk_g - 1: self.d2empty.ready(k)
            ch = chunk_g

            def op():
                assert self.sbuf[b] == ("w", g) and self.w_parts[b] == set(range(CONSUMERS)), \
                    f"GEMM 2 of stage {g} read buffer {b} in state {self.sbuf[b]} / {self.w_parts[b]}"
                assert self.stage[s] == g, f"GEMM 2 of stage {g} read stage slot {s} holding {self.stage[s]}"
                if first:
                    assert self.d2 == ("read", ch - 1) and (ch == 0 or self.d2_readers == set(range(CONSUMERS))), \
                        f"chunk {ch} started on D2 in state {self.d2} / {self.d2_readers}"
                    self.d2 = ("acc", ch)
                assert self.d2 == ("acc", ch)
                if last:
                    self.d2 = ("full", ch)
                    self.d2_readers = set()
                self.sbuf[b] = ("free", g)
                self.stage_readers[s].add("g2")
                self.stages_done += 1
            self.pipe.append(op)
            self.pipe.append(self.empty[s].arrive)
            if last:
                self.pipe.append(self.d2full.arrive)
                chunk_g += 1
            g2_g += 1
            advance(g2)
            if not g2["valid"]:
                ident = taken[g2["seq"] & 1]
                assert ident is not None, "the trailing cursor outran the leading one"
                g2["seq"] += 1
                open_(g2, ident)

        yield from take_g1()
        open_(g2, g1["item"])
        g2["seq"] = 1
        if g1["valid"]:
            yield from gemm1()
        while g2["valid"]:
            boundary = g1["valid"] and g1["ci"] == 0 and g1["st"] == 0
            if g1["valid"] and not boundary:
                yield from gemm1()
            yield from gemm2()
            if boundary:
                yield from gemm1()
            yield None

    def consumer(self, w):
        seq = g = chunk_g = item_g = 0
        while True:
            yield lambda k=seq: self.item.ready(k)
            ident, tag = self.ring[seq & 1]
            assert tag == seq, f"consumer {w} read a stale / overwritten ring entry"
            seq += 1
            if ident >= self.n_items:
                return
            if item_g > 0:
                yield lambda k=item_g - 1: self.a_free.ready(k)
            item_g += 1
            assert self.gemm1_left.get(item_g - 2, 0) == 0, "rows' image overwritten under a running GEMM 1"
            self.a_tmem_parts[w] = item_g - 1
            self.a_tmem.arrive()
            for nst in self.chunks_of(ident):
                for _ in range(nst):
                    s, b = g % STAGES, g % BUFS
                    yield lambda k=g // STAGES, s=s: self.full[s].ready(k)
                    yield lambda k=g // BUFS, b=b: self.tfull[b].ready(k)
                    assert self.stage[s] == g, f"consumer {w} read stage slot {s} holding {self.stage[s]} at stage {g}"
                    assert self.sbuf[b] in (("S", g), ("w", g)), f"consumer {w} found buffer {b} in state {self.sbuf[b]} at stage {g}"
                    yield None
                    self.w_parts[b].add(w)
                    self.sbuf[b] = ("w", g)
                    self.stage_readers[s].add(w)
                    self.wfull[b].arrive()
                    self.empty[s].arrive()
                    g += 1
                yield lambda k=chunk_g: self.d2full.ready(k)
                assert self.d2 == ("full", chunk_g) or (self.d2 == ("read", chunk_g)), f"consumer {w} read D2 in state {self.d2}, chunk {chunk_g}"
                self.d2_readers.add(w)
                if self.d2_readers == set(range(CONSUMERS)):
                    self.d2 = ("read", chunk_g)
                self.d2empty.arrive()
                chunk_g += 1

    # ---- the scheduler ---------------------------------------------------------------------------
    def run(self, max_steps=2_000_000):
        roles = {"copy": self.copy_warp(), "mma": self.mma_warp()}
        roles.update({f"c{w}": self.consumer(w) for w in range(CONSUMERS)})
        waiting = {name: None for name in roles}     # None: runnable; else the predicate it waits for
        for _ in range(max_steps):
            choices = [("role", n) for n, p in waiting.items() if p is None or p()]
            if self.tma:
                choices += [("tma", i) for i in range(len(self.tma))]
            if self.pipe:
                choices.append(("pipe", 0))
            if not choices:
                assert not roles, f"deadlock: {sorted(roles)} wait forever"
                return
            # every actor has its own speed in a run (a warp that is rarely scheduled, a slow tensor pipe, ...)
            kind, x = self.rng.choices(choices, weights=[self.speed.setdefault(c if c[0] == "role" else c[0], self.rng.choice([0.02, 1.0, 1.0, 10.0]))
                                                         for c in choices])[0]
            if kind == "tma":
                self.tma.pop(x)()
            elif kind == "pipe":
                self.pipe.pop(0)()
            else:
                try:
                    waiting[x] = roles[x].send(None)
                except StopIteration:
                    del roles[x], waiting[x]
        raise AssertionError("the model did not finish")


def run(seed, item_shapes, skip_prob=0.3, first_item_guard=True):
    m = Model(seed, item_shapes, skip_prob, first_item_guard)
    m.run()
    return m
