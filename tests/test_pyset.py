"""The oracle's CPython-set emulation against the live interpreter's `set`:
iteration order, pop order and copy layout decide which executor is moved first
(spark_sched_sim.py:714-743, executor_tracker.py:134-135)."""
import random
import sys

import pytest

from oracle import lib


class EmuSet:
    def __init__(self, h=None):
        self.L = lib()
        self.h = h if h is not None else self.L.orc_pyset_new()

    def add(self, k): self.L.orc_pyset_add(self.h, k)
    def remove(self, k): assert self.L.orc_pyset_remove(self.h, k) == 0
    def pop(self): return self.L.orc_pyset_pop(self.h)
    def copy(self): return EmuSet(self.L.orc_pyset_copy(self.h))
    def __len__(self): return self.L.orc_pyset_len(self.h)

    def list(self):
        import ctypes as C
        buf = (C.c_int32 * 4096)()
        n = self.L.orc_pyset_list(self.h, buf)
        return list(buf[:n])

    def __del__(self):
        self.L.orc_pyset_free(self.h)


@pytest.mark.skipif(sys.version_info[:2] != (3, 12), reason="layout pinned to CPython 3.12")
def test_reward_set_ascending_shortcut():
    """The kernel's reward sum (ssb_sim.cuh compute_jobtime) iterates set(active_old + active_new)
    in ascending order whenever max(id) < F(n), F = 8/32/128/512/2048 for n < 5/19/77/307/1229
    distinct ids inserted in ascending order.  Check that claim against the interpreter."""
    def final_size(n):
        return 8 if n < 5 else 32 if n < 19 else 128 if n < 77 else 512 if n < 307 else 2048 if n < 1229 else 0

    rnd = random.Random(7)
    used = 0
    for _ in range(20000):
        J = rnd.choice([8, 50, 200, 1000])
        pool = sorted(rnd.sample(range(J), rnd.randint(0, min(J, 320))))
        cut = rnd.randint(0, len(pool))
        old = pool[:cut]
        new = [x for x in old if rnd.random() > 0.1] + pool[cut:][:rnd.randint(0, 5)]
        s = set(old + new)
        ids = sorted(s)
        if ids and ids[-1] < final_size(len(ids)):
            used += 1
            assert list(s) == ids
    assert used > 5000


def _register_table_order(ids):
    """Python mirror of the kernel's in-register emulation for <= 18 ascending distinct ids < 255
    (ssb_sim.cuh compute_jobtime): 8-slot table without linear probes, rebuilt into a 32-slot table at
    the 5th element, where the 9-slot linear probe is a find-first-zero on the occupancy mask."""
    occ8, t8, occ, T, cnt, big = 0, {}, 0, {}, 0, False

    def put32(key):
        nonlocal occ
        perturb, i = key, key & 31
        while True:
            if not (occ >> i) & 1:
                break
            if i + 9 <= 31:
                m = (~occ >> (i + 1)) & 0x1FF
                if m:
                    i = i + (m & -m).bit_length()
                    break
            perturb >>= 5
            i = (i * 5 + 1 + perturb) & 31
        occ |= 1 << i
        T[i] = key

    for key in ids:
        if big:
            put32(key); cnt += 1
            continue
        perturb, i = key, key & 7
        while (occ8 >> i) & 1:
            perturb >>= 5
            i = (i * 5 + 1 + perturb) & 7
        occ8 |= 1 << i
        t8[i] = key
        cnt += 1
        if cnt == 5:
            for s in range(8):
                if (occ8 >> s) & 1:
                    put32(t8[s])
            big = True
    return [T[s] for s in range(32) if (occ >> s) & 1] if big else [t8[s] for s in range(8) if (occ8 >> s) & 1]


@pytest.mark.skipif(sys.version_info[:2] != (3, 12), reason="layout pinned to CPython 3.12")
def test_reward_small_set_register_emulation():
    rnd = random.Random(11)
    for _ in range(30000):
        n = rnd.randint(1, 18)
        ids = sorted(rnd.sample(range(rnd.choice([20, 50, 200, 255])), n))
        dup = [x for x in ids if rnd.random() < 0.7]  # second list repeats part of the first
        assert list(set(ids + dup)) == _register_table_order(ids), ids


@pytest.mark.skipif(sys.version_info[:2] != (3, 12), reason="layout pinned to CPython 3.12")
@pytest.mark.parametrize("universe", [3, 10, 50, 100, 200])
def test_pyset_matches_cpython(universe):
    rnd = random.Random(universe)
    for trial in range(200):
        real, emu = set(), EmuSet()
        if rnd.random() < 0.5:
            for i in range(universe):  # set(range(n)) as executor_tracker.py:41
                real.add(i); emu.add(i)
        for _ in range(rnd.randrange(1, 120)):
            op = rnd.random()
            if op < 0.45:
                k = rnd.randrange(universe)
                real.add(k); emu.add(k)
            elif op < 0.75 and real:
                k = rnd.choice(sorted(real))
                real.remove(k); emu.remove(k)
            elif op < 0.85 and real:
                assert real.pop() == emu.pop()
            elif op < 0.95:
                real, emu = real.copy(), emu.copy()
            else:  # set(generator) over a copy with a filter (spark_sched_sim.py:720-726)
                keep = rnd.randrange(2, 5)
                real = set(x for x in real.copy() if x % keep)
                src = emu.copy().list()
                emu = EmuSet()
                for x in src:
                    if x % keep:
                        emu.add(x)
            assert list(real) == emu.list(), (universe, trial)
            assert len(real) == len(emu)


def test_small_set_slot_order():
    """Claim used by the reward kernel (compute_jobtime_w): a set built from fewer than 19 ascending distinct
    ints whose values are distinct modulo the final table size (8 below five elements, else 32) iterates in
    slot order, i.e. sorted by value & (size - 1) -- checked against the interpreter."""
    import random

    rnd = random.Random(5)
    checked = 0
    for _ in range(60000):
        n = rnd.randint(1, 18)
        keys = sorted(rnd.sample(range(0, 256), n))
        mask = 7 if n < 5 else 31
        if len({k & mask for k in keys}) != n:
            continue
        assert list(set(keys)) == sorted(keys, key=lambda k: k & mask), keys
        checked += 1
    assert checked > 10000
