"""CPU: discrete-event check of the mbarrier protocol of the CTA-pair GEMM (csrc/gemm_tc2.cu) -- no GPU needed.

Models the roles of both CTAs (TMA producer, 4 splitter warps, MMA issuer in the leader, 8 epilogue warps) with
mbarrier semantics (arrival count per phase, parity waits) and random role speeds, and checks that every run
terminates (no dead-lock), that no barrier ever receives more arrivals than its count in one phase, that a stage is
never refilled before the MMAs that read it were committed, and that an accumulator is never overwritten before
both epilogues drained it.  `python tools/tc2_protocol_sim.py` runs 300 random schedules.
"""
import random

STAGES, ACC = 4, 2


class MBar:
    def __init__(self, count, name):
        self.count, self.pending, self.phase, self.name = count, count, 0, name

    def arrive(self):
        assert self.pending > 0, f"{self.name}: more arrivals than its count in one phase"
        self.pending -= 1
        if self.pending == 0:
            self.phase ^= 1
            self.pending = self.count

    def ready(self, parity):          # mbarrier.try_wait.parity: true once the phase with this parity completed
        return self.phase != parity


def simulate(items, kblocks, seed):
    rng = random.Random(seed)
    # per CTA: full_raw / empty per stage, acc_full per accumulator (local); leader only: full_split, acc_empty
    full_raw = [[MBar(1, f"full_raw{c}.{s}") for s in range(STAGES)] for c in range(2)]
    empty = [[MBar(1, f"empty{c}.{s}") for s in range(STAGES)] for c in range(2)]
    acc_full = [[MBar(1, f"acc_full{c}.{a}") for a in range(ACC)] for c in range(2)]
    full_split = [MBar(8, f"full_split.{s}") for s in range(STAGES)]          # 4 splitter warps x 2 CTAs
    acc_empty = [MBar(16, f"acc_empty.{a}") for a in range(ACC)]              # 8 epilogue warps x 2 CTAs
    stage_busy = [[False] * STAGES for _ in range(2)]     # filled and not yet released by a commit
    acc_busy = [[0] * ACC for _ in range(2)]              # epilogue warps of this CTA still to drain it

    def tma(c):
        stage, phase = 0, 0
        for _ in range(items):
            for _ in range(kblocks):
                while not empty[c][stage].ready(phase ^ 1):
                    yield
                assert not stage_busy[c][stage], "stage refilled before its MMAs were committed"
                stage_busy[c][stage] = True
                for _ in range(rng.randint(0, 3)):
                    yield                                   # load latency
                full_raw[c][stage].arrive()                 # complete_tx
                stage = (stage + 1) % STAGES
                phase ^= stage == 0

    def splitter(c, w):
        stage, phase = 0, 0
        for _ in range(items):
            for _ in range(kblocks):
                while not full_raw[c][stage].ready(phase):
                    yield
                for _ in range(rng.randint(0, 2)):
                    yield
                full_split[stage].arrive()                  # remote arrive on the leader
                stage = (stage + 1) % STAGES
                phase ^= stage == 0

    def mma():
        stage, phase = 0, 0
        for it in range(items):
            acc, acc_phase = it & 1, (it >> 1) & 1
            while not acc_empty[acc].ready(acc_phase ^ 1):
                yield
            assert acc_busy[0][acc] == 0 and acc_busy[1][acc] == 0, "accumulator overwritten before it was drained"
            for kb in range(kblocks):
                while not full_split[stage].ready(phase):
                    yield
                assert stage_busy[0][stage] and stage_busy[1][stage], "MMA on a stage that is not filled"
                for _ in range(rng.randint(0, 2)):
                    yield                                   # MMAs execute
                for c in range(2):                          # tcgen05.commit multicast
                    stage_busy[c][stage] = False
                    empty[c][stage].arrive()
                if kb == kblocks - 1:
                    for c in range(2):
                        acc_busy[c][acc] = 8
                        acc_full[c][acc].arrive()
                stage = (stage + 1) % STAGES
                phase ^= stage == 0

    def epilogue(c, w):
        for it in range(items):
            acc, acc_phase = it & 1, (it >> 1) & 1
            while not acc_full[c][acc].ready(acc_phase):
                yield
            for _ in range(rng.randint(0, 6)):
                yield                                       # drain + stores
            acc_busy[c][acc] -= 1
            acc_empty[acc].arrive()                         # remote arrive on the leader

    procs = [tma(0), tma(1), mma()] + [splitter(c, w) for c in range(2) for w in range(4)] + \
            [epilogue(c, w) for c in range(2) for w in range(8)]
    live = list(procs)
    idle_rounds = 0
    while live:
        progressed = False
        rng.shuffle(live)
        for p in list(live):
            before = snapshot(full_raw, empty, acc_full, full_split, acc_empty)
            try:
                next(p)
            except StopIteration:
                live.remove(p)
                progressed = True
                continue
            if snapshot(full_raw, empty, acc_full, full_split, acc_empty) != before:
                progressed = True
        idle_rounds = 0 if progressed else idle_rounds + 1
        assert idle_rounds < 50, f"dead-lock: {len(live)} roles blocked (items={items}, kblocks={kblocks}, seed={seed})"
    return True


def snapshot(*groups):
    out = []
    for g in groups:
        for b in (x for row in g for x in (row if isinstance(row, list) else [row])):
            out.append((b.phase, b.pending))
    return tuple(out)


if __name__ == "__main__":
    n = 0
    for seed in range(300):
        r = random.Random(1000 + seed)
        simulate(items=r.randint(1, 7), kblocks=r.randint(1, 9), seed=seed)
        n += 1
    print(f"{n} random schedules: no dead-lock, no over-arrival, no stage / accumulator hazard")
