"""GPU parity: the CUDA path, called through the C ABI, against (1) the committed reference traces
and (2) the CPU oracle on seeded inputs.  Bit-exact: event order, executor assignments,
observations, f64 wall/completion times; rewards bit-exact for beta == 0 and 1e-12 relative for
beta > 0 (exp())."""
import numpy as np
import pytest

from helpers import bank_for, golden_names, load_golden, replay_and_compare

pytestmark = pytest.mark.gpu


def env_cfg_of(tr):
    return {"num_executors": tr["num_executors"],
            "job_arrival_cap": tr["job_arrival_cap"] if tr["job_arrival_cap"] > 0 else None,
            "job_arrival_rate": tr["job_arrival_rate"], "moving_delay": tr["moving_delay"],
            "warmup_delay": tr["warmup_delay"], "beta": tr["beta"]}


class CudaAdapter:
    """Presents one slot of a BatchedSparkSchedSimEnv with the OracleEnv interface; every step goes
    through ssb_step_host (host buffers in, observation headers out)."""

    def __init__(self, bank, tr, B=3, slot=1, max_events=0):
        from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

        self.B, self.slot, self.tr, self.max_events = B, slot, tr, max_events
        self.env = BatchedSparkSchedSimEnv(
            env_cfg_of(tr), num_envs=B, bank=bank, max_jobs=len(tr["job_template"]) + 4,
            tape_capacity=len(tr["tape"]) + 8, log_capacity=int(tr["ev_count"][-1]) + 64,
            history_capacity=(int(tr["hist_ptr"][-1]) + 8) if "hist_ptr" in tr else 0)
        self.hdr = None

    def reset_trace(self, ta, tm, tape):
        for b in range(self.B):
            self.env.load_trace(b, ta, tm, tape)
        self.hdr = self.env.reset_host(np.full(self.B, self.tr["seed"], np.uint64)).copy()
        assert (self.hdr["error"] == 0).all(), self.hdr["error"]
        return self.obs()

    def reset_seed(self, seed, time_limit=np.inf):
        self.hdr = self.env.reset_host(np.full(self.B, seed, np.uint64),
                                       np.full(self.B, time_limit, np.float64)).copy()
        assert (self.hdr["error"] == 0).all(), self.hdr["error"]
        return self.obs()

    def step(self, a, n):
        self.hdr = self.env.step_host(np.full(self.B, a, np.int32), np.full(self.B, n, np.int32),
                                      max_events=self.max_events).copy()
        calls = 1
        while self.hdr[self.slot]["pending"]:  # budgeted mode: continue; the action is ignored now
            self.hdr = self.env.step_host(np.full(self.B, -7, np.int32), np.full(self.B, -7, np.int32),
                                          max_events=self.max_events).copy()
            calls += 1
            assert calls < 100000
        h = self.hdr[self.slot]
        return int(h["error"]), float(h["reward"]), bool(h["terminated"])

    def obs(self):
        return self.env.obs(self.slot, self.hdr)

    @property
    def wall_time(self):
        return float(self.hdr[self.slot]["wall_time"])

    def log_size(self):
        return self.env.log_size(self.slot)

    def log(self):
        return self.env.log(self.slot)

    def job_times(self):
        return self.env.jobs(self.slot)

    def history(self):
        return self.env.history(self.slot)

    def fair_action(self, dynamic_partition):
        a, n = self.env.fair_actions(dynamic_partition)
        return int(a[self.slot].item()), int(n[self.slot].item())


@pytest.mark.parametrize("name", golden_names())
def test_cuda_replays_reference_tape(name):
    tr = load_golden(name)
    env = CudaAdapter(bank_for(tr), tr)
    replay_and_compare(env, tr, "tape", check_policy=env.fair_action,
                       reward_rtol=1e-12 if tr["beta"] > 0 else 0.0)


@pytest.mark.parametrize("name", [n for n in golden_names() if "philox" in n])
def test_cuda_replays_reference_from_seed(name):
    """On-device Philox sampling of jobs and durations against the Philox-plugged reference run."""
    tr = load_golden(name)
    env = CudaAdapter(bank_for(tr), tr)
    replay_and_compare(env, tr, "seed", check_policy=env.fair_action,
                       reward_rtol=1e-12 if tr["beta"] > 0 else 0.0)


def test_cuda_invalid_actions(bank):
    tr = load_golden("e10_j8_fair_s2_philox")
    env = CudaAdapter(bank, tr)
    obs = env.reset_seed(tr["seed"])
    N = obs["nodes"].shape[0]
    nsched = int(obs["nodes"][:, 2].sum())
    wall0, h0 = env.wall_time, env.hdr.copy()
    assert env.step(N, 1)[0] == 1
    assert env.step(0, 0)[0] == 1
    assert env.step(0, 11)[0] == 1
    if nsched < N:
        assert env.step(nsched, 1)[0] == 2
    # rejected actions leave the state untouched: the recorded episode still replays exactly
    replay_and_compare(CudaAdapter(bank, tr), tr, "seed")
    assert env.wall_time == wall0
    rc, _, _ = env.step(int(tr["actions"][0][0]), int(tr["actions"][0][1]))
    assert rc == 0


@pytest.mark.parametrize("E,J,policy", [(10, 50, "fair"), (10, 20, "fifo"), (50, 30, "fair"), (3, 10, "fair"),
                                        # 32 < E <= 64: two executor slots per lane in the batched fast path
                                        (33, 20, "fair"), (64, 40, "fifo"), (50, 200, "fair"),
                                        # E > 64: general path only
                                        (100, 30, "fair")])
def test_fused_rollout_matches_oracle(bank, E, J, policy):
    """Fused on-device policy+step episodes vs the CPU oracle from the same seeds: every job's
    arrival/completion time, the final wall time, and the decision/event counts."""
    from oracle import OracleEnv
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    B = 16 if J < 200 else 6  # (50, 200) = BASELINE config 4's episode shape
    cfg = {"num_executors": E, "job_arrival_cap": J, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank)
    seeds = np.arange(1000, 1000 + B, dtype=np.uint64)
    env.reset_host(seeds)
    env.rollout_fair(1_000_000, dynamic_partition=(policy == "fair"), auto_reset=False)
    hdr = env.hdr()
    assert (hdr["error"] == 0).all(), hdr["error"]
    assert (hdr["terminated"] == 1).all()
    tot_dec = tot_ev = 0
    for b in range(B):
        orc = OracleEnv(bank, E, J, 2000.0, 1000.0, 4.0e-5)
        dec, ev = orc.run_fair_episode(int(seeds[b]), policy == "fair")
        tot_dec += dec
        tot_ev += ev
        ta, tc, tm = orc.job_times()
        gta, gtc, gtm = env.jobs(b)
        assert np.array_equal(tm, gtm) and np.array_equal(ta, gta), b
        assert np.array_equal(tc, gtc), (b, "job completion times")
        assert hdr["wall_time"][b] == orc.wall_time
    st = env.stats()
    assert st["decisions"] == tot_dec and st["events"] == tot_ev
    assert st["episodes"] == B


def test_full_size_batch_properties(bank):
    """BASELINE config 2 at full size (4096 envs, 50 jobs x 10 executors): no env errors, every
    episode terminates, envs sharing a seed are identical, a sample matches the oracle, and the
    run is deterministic."""
    from oracle import OracleEnv
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    B = 4096
    cfg = {"num_executors": 10, "job_arrival_cap": 50, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank)
    seeds = (1234 + np.arange(B) // 2).astype(np.uint64)  # pairs share a job sequence
    walls = []
    for rep in range(2):
        env.reset_stats()
        env.reset_host(seeds)
        env.rollout_fair(1_000_000, True, auto_reset=False)
        hdr = env.hdr()
        assert (hdr["error"] == 0).all()
        assert (hdr["terminated"] == 1).all()
        walls.append(hdr["wall_time"].copy())
    assert np.array_equal(walls[0], walls[1])
    assert np.array_equal(walls[0][0::2], walls[0][1::2])
    st = env.stats()
    assert st["episodes"] == B and st["decisions"] > 400 * B
    for b in (0, 1777, 4095):
        orc = OracleEnv(bank, 10, 50, 2000.0, 1000.0, 4.0e-5)
        orc.run_fair_episode(int(seeds[b]), True)
        assert np.array_equal(orc.job_times()[1], env.jobs(b)[1])


def test_auto_reset_rollout(bank):
    """auto_reset: a finished env re-seeds itself with seed + seed_step * reset_count
    (rollout_worker.py:118-120) and keeps deciding; decisions per env are exact."""
    from oracle import OracleEnv
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    B, K = 8, 700
    cfg = {"num_executors": 10, "job_arrival_cap": 8, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank)
    seeds = np.arange(50, 50 + B, dtype=np.uint64)
    env.reset_host(seeds)
    env.rollout_fair(K, True, auto_reset=True, seed_step=100)
    st = env.stats()
    assert st["decisions"] == B * K
    hdr = env.hdr()
    assert (hdr["error"] == 0).all()
    # replay env 3 on the oracle: K decisions across episodes seeded 53, 153, 253, ...
    orc = OracleEnv(bank, 10, 8, 2000.0, 1000.0, 4.0e-5)
    k, ep = 0, 0
    while k < K:
        orc.reset_seed(53 + 100 * ep)
        term = False
        while not term and k < K:
            a, n = orc.fair_action(True)
            rc, _, term = orc.step(a, n)
            assert rc == 0
            k += 1
        ep += 1
    assert hdr["wall_time"][3] == orc.wall_time
    o, g = orc.obs(), env.obs(3, hdr)
    for key in ("nodes", "edge_links", "dag_ptr", "exec_supplies"):
        assert np.array_equal(o[key], g[key]), key


def test_fused_rollout_fuzz_vs_oracle(bank):
    """Seeded sweep over the configuration space (executors 1..64, job caps, delays, fair/FIFO): whole episodes of
    the fused rollout equal the oracle's in every job completion time, final wall time and decision/event count."""
    from oracle import OracleEnv
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    rng = np.random.default_rng(7)
    for case in range(14):
        E = int(rng.choice([1, 2, 5, 10, 17, 31, 32, 33, 40, 50, 63, 64]))
        J = int(rng.integers(1, 26))
        fair = bool(rng.integers(0, 2))
        moving, warm = float(rng.choice([0.0, 500.0, 2000.0])), float(rng.choice([0.0, 1000.0, 3000.0]))
        rate = float(rng.choice([1e-5, 4e-5, 2e-4]))
        B = 4
        cfg = {"num_executors": E, "job_arrival_cap": J, "job_arrival_rate": rate,
               "moving_delay": moving, "warmup_delay": warm}
        env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank)
        seeds = (rng.integers(0, 2**31, B)).astype(np.uint64)
        env.reset_host(seeds)
        env.rollout_fair(1_000_000, dynamic_partition=fair, auto_reset=False)
        hdr = env.hdr()
        assert (hdr["error"] == 0).all() and (hdr["terminated"] == 1).all(), (case, cfg, hdr["error"])
        dec = ev = 0
        for b in range(B):
            orc = OracleEnv(bank, E, J, moving, warm, rate)
            d, e = orc.run_fair_episode(int(seeds[b]), fair)
            dec += d; ev += e
            assert np.array_equal(orc.job_times()[1], env.jobs(b)[1]), (case, cfg, b)
            assert hdr["wall_time"][b] == orc.wall_time, (case, cfg, b)
        st = env.stats()
        assert (st["decisions"], st["events"]) == (dec, ev), (case, cfg)
        del env


def test_rollout_transitions_match_oracle(bank):
    """ssb_rollout_fair_traj: the recorded (wall_time, action, reward, flags) rows are what a host loop
    over the oracle's reset()/fair_action()/step() produces, across auto-resets."""
    import torch
    from oracle import OracleEnv
    from spark_sched_sim_b200 import _native as nat
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    B, K = 6, 900
    cfg = {"num_executors": 10, "job_arrival_cap": 8, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank)
    seeds = np.arange(500, 500 + B, dtype=np.uint64)
    env.reset_host(seeds)
    host = torch.empty(B * K * nat.TRANSITION_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    tr = env.rollout_fair_traj(K, True, auto_reset=True, seed_step=1000, host=host)
    assert tr.shape == (B, K) and (env.hdr()["error"] == 0).all()
    for b in (0, 5):
        orc = OracleEnv(bank, 10, 8, 2000.0, 1000.0, 4.0e-5)
        k = ep = 0
        while k < K:
            orc.reset_seed(int(seeds[b]) + 1000 * ep)
            term, first = False, ep > 0
            while not term and k < K:
                wall0 = orc.wall_time
                a, n = orc.fair_action(True)
                rc, rew, term = orc.step(a, n)
                assert rc == 0
                row = tr[b, k]
                assert (row["wall_time"], row["stage_idx"], row["num_exec"]) == (wall0, a, n), (b, k)
                assert row["reward"] == rew and (row["flags"] & 1) == int(term), (b, k)
                assert bool(row["flags"] & 4) == first, (b, k)
                first = False
                k += 1
            ep += 1
        assert ep > 1


def _oracle_transitions(bank, cfg, seed, seed_step, K, time_limit=np.inf, beta=0.0, mean_time_limit=0.0):
    """K decisions of the oracle under the fair policy with the rollout worker's re-seeding rule.
    mean_time_limit > 0: every episode's limit is the StochasticTimeLimit draw of its seed."""
    from oracle import OracleEnv
    from philox_ref import time_limit_draw

    orc = OracleEnv(bank, cfg["num_executors"], cfg["job_arrival_cap"], cfg["moving_delay"], cfg["warmup_delay"],
                    cfg["job_arrival_rate"], beta=beta)
    rows, k, ep = [], 0, 0
    while k < K:
        if mean_time_limit > 0:
            time_limit = time_limit_draw(int(seed) + seed_step * ep, mean_time_limit)
        orc.reset_seed(int(seed) + seed_step * ep, time_limit)
        done = False
        while not done and k < K:
            wall0 = orc.wall_time
            a, n = orc.fair_action(True)
            rc, rew, term = orc.step(a, n)
            assert rc == 0
            trunc = orc.wall_time >= time_limit  # StochasticTimeLimit (wrappers/stochastic_time_limit.py:29-30)
            rows.append((wall0, rew, a, n, int(term), int(trunc)))
            done = term or trunc
            k += 1
        ep += 1
    return rows, ep


def test_rollout_with_time_limit_truncation(bank):
    """Continuous arrivals up to a time limit (no job cap): episodes end by truncation, the fused rollout
    re-seeds them, and every transition equals the oracle's."""
    import torch
    from spark_sched_sim_b200 import _native as nat
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    B, K, TL = 4, 500, 4.0e5
    cfg = {"num_executors": 10, "job_arrival_cap": 0, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank, max_jobs=96)
    seeds = np.arange(900, 900 + B, dtype=np.uint64)
    env.reset_host(seeds, time_limits=np.full(B, TL))
    host = torch.empty(B * K * nat.TRANSITION_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    tr = env.rollout_fair_traj(K, True, auto_reset=True, seed_step=50, host=host)
    assert (env.hdr()["error"] == 0).all()
    n_trunc = 0
    for b in range(B):
        rows, eps = _oracle_transitions(bank, cfg, seeds[b], 50, K, time_limit=TL)
        for k, (wall0, rew, a, n, term, trunc) in enumerate(rows):
            r = tr[b, k]
            assert (r["wall_time"], r["reward"], r["stage_idx"], r["num_exec"]) == (wall0, rew, a, n), (b, k)
            assert (r["flags"] & 1, (r["flags"] >> 1) & 1) == (term, trunc), (b, k)
            n_trunc += trunc
        assert eps > 1
    assert n_trunc > 0


def test_capacity_is_reported_and_the_handle_grows(bank):
    """Time-limited arrivals without a job cap on a handle that is too small: the reset reports SSB_ENV_CAPACITY
    (nothing is clipped), reset_host(grow=True) re-creates the handle with twice the room until the episodes fit, and
    the grown handle's rollout equals the oracle's -- i.e. what a large enough handle gives."""
    import torch
    from spark_sched_sim_b200 import _native as nat
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    B, K, TL = 4, 200, 4.0e5
    cfg = {"num_executors": 10, "job_arrival_cap": 0, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank, max_jobs=4)
    env.set_autoreset(True, 50)
    seeds = np.arange(900, 900 + B, dtype=np.uint64)
    hdr = env.reset_host(seeds, time_limits=np.full(B, TL))
    assert (hdr["error"] == nat.ENV_CAPACITY).any()
    hdr = env.reset_host(seeds, time_limits=np.full(B, TL), grow=True)
    assert (hdr["error"] == 0).all() and env.max_jobs > 4
    assert env._settings["autoreset"] == (True, 50)
    host = torch.empty(B * K * nat.TRANSITION_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    # a grown handle can still be too small for a LATER episode of the rollout: size it for the tail first
    need = BatchedSparkSchedSimEnv.required_job_capacity(cfg["job_arrival_rate"], TL, tail=1e-12)
    if env.max_jobs < need:
        env.grow(need)
        env.reset_host(seeds, time_limits=np.full(B, TL))
    tr = env.rollout_fair_traj(K, True, auto_reset=True, seed_step=50, host=host)
    assert (env.hdr()["error"] == 0).all()
    for b in range(B):
        rows, _ = _oracle_transitions(bank, cfg, seeds[b], 50, K, time_limit=TL)
        for k, (wall0, rew, a, n, term, trunc) in enumerate(rows):
            r = tr[b, k]
            assert (r["wall_time"], r["reward"], r["stage_idx"], r["num_exec"]) == (wall0, rew, a, n), (b, k)
    # the geometric tail: with rate * limit = 16 jobs on average, 1e-12 is a few hundred jobs
    assert 100 < need < 1000


def test_step_api_autoreset_matches_oracle(bank):
    """ssb_set_autoreset: a step() on a finished env re-seeds it (seed + seed_step * reset_count), ignores the
    action and flags was_reset; the sequence of real transitions equals the oracle loop with explicit resets."""
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    B, K = 4, 700
    cfg = {"num_executors": 10, "job_arrival_cap": 6, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank)
    seeds = np.arange(300, 300 + B, dtype=np.uint64)
    env.reset_host(seeds)
    env.set_autoreset(True, 11)
    got = [[] for _ in range(B)]
    n_resets = 0
    while min(len(g) for g in got) < K:
        wall0 = env.hdr()["wall_time"].copy()
        a, n = env.fair_actions(True)
        a_h, n_h = a.cpu().numpy(), n.cpu().numpy()
        h = env.step_host(a_h, n_h).copy()
        assert (h["error"] == 0).all()
        for b in range(B):
            if h["was_reset"][b]:
                assert h["reward"][b] == 0.0 and h["wall_time"][b] == 0.0 and not h["terminated"][b]
                n_resets += 1
            else:
                got[b].append((wall0[b], h["reward"][b], int(a_h[b]), int(n_h[b]), int(h["terminated"][b]), 0))
    assert n_resets > B
    for b in range(B):
        rows, _ = _oracle_transitions(bank, cfg, seeds[b], 11, K)
        assert got[b][:K] == rows, b


def test_step_fair_host_matches_oracle(bank):
    """ssb_step_fair_host: the step kernel's own fair-scheduler suggestion drives the next call (one call and one
    synchronisation per decision); with auto-reset the real transitions equal the oracle loop's."""
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    B, K = 4, 600
    cfg = {"num_executors": 10, "job_arrival_cap": 6, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank)
    seeds = np.arange(700, 700 + B, dtype=np.uint64)
    env.reset_host(seeds)
    env.set_autoreset(True, 13)
    a0, n0 = env.fair_actions(True)
    a, n = a0.cpu().numpy().copy(), n0.cpu().numpy().copy()
    na, nn = np.zeros(B, np.int32), np.zeros(B, np.int32)
    got = [[] for _ in range(B)]
    wall = env.hdr()["wall_time"].copy()
    while min(len(g) for g in got) < K:
        h = env.step_fair_host(a, n, na, nn, True).copy()
        assert (h["error"] == 0).all()
        for b in range(B):
            if not h["was_reset"][b]:
                got[b].append((wall[b], h["reward"][b], int(a[b]), int(n[b]), int(h["terminated"][b]), 0))
        wall = h["wall_time"].copy()
        a, n = na.copy(), nn.copy()
    for b in range(B):
        rows, _ = _oracle_transitions(bank, cfg, seeds[b], 13, K)
        assert got[b][:K] == rows, b


def test_collect_stats_matches_host_metrics(bank):
    """ssb_collect_stats sums == the reference's metrics (spark_sched_sim/metrics.py) evaluated per env on the
    host from the per-job times, mid-episode."""
    from spark_sched_sim_b200 import parallel
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    B = 70
    cfg = {"num_executors": 10, "job_arrival_cap": 12, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank)
    env.reset_host(np.arange(B, dtype=np.uint64) + 3)
    env.rollout_fair(90, True, auto_reset=False)
    v = env.collect_stats().cpu().numpy()
    hdr = env.hdr()
    want = np.zeros(6)
    for b in range(B):
        ta, tc, tm, state = env.jobs(b, with_state=True)
        wall = hdr["wall_time"][b]
        arrived = state != 0
        jt = (np.minimum(tc[arrived], wall) - ta[arrived]).sum()
        if wall > 0:
            want[0] += jt / wall; want[1] += 1
        done = state == 2
        want[2] += done.sum(); want[3] += arrived.sum(); want[4] += (tc[done] - ta[done]).sum(); want[5] += wall
    assert np.allclose(v[:6], want, rtol=1e-12), (v, want)
    s = parallel.stats_from_sums(v)
    assert 0 < s["avg_num_jobs"] < 12 and s["num_job_arrivals"] <= 12


def test_stochastic_time_limit_on_device(bank):
    """ssb_set_mean_time_limit: resets without an explicit limit and auto-resets draw the episode's limit
    ~ Exp(mean) from the seed's Philox LIMIT stream (StochasticTimeLimit); the fused rollout then equals the
    oracle driven with the spec's draws (oracle/philox_ref.py:time_limit_draw)."""
    import torch
    from spark_sched_sim_b200 import _native as nat
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    B, K, MEAN = 4, 500, 3.0e5
    cfg = {"num_executors": 10, "job_arrival_cap": 0, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank, max_jobs=128)
    env.set_mean_time_limit(MEAN)
    seeds = np.arange(60, 60 + B, dtype=np.uint64)
    env.reset_host(seeds)
    host = torch.empty(B * K * nat.TRANSITION_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    tr = env.rollout_fair_traj(K, True, auto_reset=True, seed_step=17, host=host)
    assert (env.hdr()["error"] == 0).all()
    n_trunc = 0
    for b in range(B):
        rows, eps = _oracle_transitions(bank, cfg, seeds[b], 17, K, mean_time_limit=MEAN)
        assert eps > 2
        for k, (wall0, rew, a, n, term, trunc) in enumerate(rows):
            r = tr[b, k]
            assert (r["wall_time"], r["reward"], r["stage_idx"], r["num_exec"]) == (wall0, rew, a, n), (b, k)
            assert (r["flags"] & 1, (r["flags"] >> 1) & 1) == (term, trunc), (b, k)
            n_trunc += trunc
    assert n_trunc > B


def test_async_rollouts_match_reference_loop(bank):
    """ssb_rollout_fair_async == RolloutWorkerAsync.collect_rollout (rollout_worker.py:160-206) restated over the
    oracle: fixed simulated duration per call, resets inside the rollout, time axis = accumulated time, state carried
    over to the next call."""
    from oracle import OracleEnv
    from spark_sched_sim_b200 import _native as nat
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    B, KMAX, DUR, STEP = 4, 2000, 2.0e6, 9
    cfg = {"num_executors": 10, "job_arrival_cap": 5, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank)
    seeds = np.arange(80, 80 + B, dtype=np.uint64)
    env.reset_host(seeds)
    calls = []
    for _ in range(3):
        traj, num, el = env.rollout_fair_async(KMAX, DUR, True, STEP)
        calls.append((traj.cpu().numpy().view(nat.TRANSITION_DTYPE).reshape(B, KMAX), num.cpu().numpy(),
                      el.cpu().numpy()))
    assert (env.hdr()["error"] == 0).all()
    n_resets = 0
    for b in range(B):
        orc = OracleEnv(bank, 10, 5, 2000.0, 1000.0, 4.0e-5)
        resets = 0
        orc.reset_seed(int(seeds[b])); resets += 1
        next_wall = 0.0
        for tr, num, el in calls:  # one collect_rollout() each
            elapsed, step = 0.0, 0
            while elapsed < DUR and step < KMAX:
                wall = next_wall
                a, n = orc.fair_action(True)
                rc, rew, term = orc.step(a, n)
                assert rc == 0
                next_wall = orc.wall_time
                r = tr[b, step]
                assert (r["wall_time"], r["stage_idx"], r["num_exec"], r["reward"]) == (elapsed, a, n, rew), (b, step)
                assert (r["flags"] & 1) == int(term), (b, step)
                elapsed += next_wall - wall
                if term:
                    orc.reset_seed(int(seeds[b]) + STEP * resets); resets += 1
                    next_wall = 0.0
                    n_resets += 1
                step += 1
            assert num[b] == step and el[b] == elapsed, (b, num[b], step)
    assert n_resets >= B


def test_rollout_with_discounted_reward(bank):
    """beta > 0: the continuously discounted reward (:866-869) goes through exp(); transitions match the
    oracle with rewards at 1e-12 relative (device exp vs libm), everything else exactly."""
    import torch
    from spark_sched_sim_b200 import _native as nat
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    B, K, beta = 3, 400, 5.0e-3
    cfg = {"num_executors": 10, "job_arrival_cap": 12, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0, "beta": beta}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank)
    seeds = np.arange(40, 40 + B, dtype=np.uint64)
    env.reset_host(seeds)
    host = torch.empty(B * K * nat.TRANSITION_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    tr = env.rollout_fair_traj(K, True, auto_reset=True, seed_step=7, host=host)
    assert (env.hdr()["error"] == 0).all()
    for b in range(B):
        rows, _ = _oracle_transitions(bank, cfg, seeds[b], 7, K, beta=beta)
        for k, (wall0, rew, a, n, term, trunc) in enumerate(rows):
            r = tr[b, k]
            assert (r["wall_time"], r["stage_idx"], r["num_exec"], r["flags"] & 1) == (wall0, a, n, term), (b, k)
            assert abs(r["reward"] - rew) <= 1e-12 * max(1.0, abs(rew)), (b, k, r["reward"], rew)


@pytest.mark.parametrize("name,budget", [("e10_j8_random_s5_philox", 1), ("e10_j8_fair_s2_philox", 7),
                                         ("e50_j8_random_s8_philox", 33), ("c2_fair_s1234_philox", 64)])
def test_budgeted_step_is_equivalent(bank, name, budget):
    """max_events > 0 (asynchronous-vector-env mode) only changes how the work is sliced over calls:
    the trajectory is bit-identical to the reference trace."""
    tr = load_golden(name)
    env = CudaAdapter(bank, tr, max_events=budget)
    replay_and_compare(env, tr, "seed", reward_rtol=1e-12 if tr["beta"] > 0 else 0.0)


@pytest.mark.parametrize("name", ["e10_j8_fair_s2_philox", "e10_j8_fifo_s3_philox", "e50_j8_fair_s7_philox",
                                  "c2_fair_s1234_philox"])
def test_facade_runs_like_examples_py(bank, name):
    """examples.py:84-102 with the drop-in classes: gym-style env + host RoundRobinScheduler reproduce
    the reference run (same actions from the same observations, same rewards, same avg job duration)."""
    from spark_sched_sim_b200 import metrics
    from spark_sched_sim_b200.env import SparkSchedSimEnv
    from spark_sched_sim_b200.schedulers import RoundRobinScheduler

    tr = load_golden(name)
    cfg = env_cfg_of(tr)
    cfg["data_sampler_cls"] = "TPCHDataSampler"
    env = SparkSchedSimEnv(cfg, bank=bank, history_capacity=int(tr["hist_ptr"][-1]) + 8)
    scheduler = RoundRobinScheduler(cfg["num_executors"], dynamic_partition=(tr["policy"] == "fair"))
    obs, _ = env.reset(seed=tr["seed"], options=None)
    terminated = truncated = False
    k = 0
    while not (terminated or truncated):
        action, _ = scheduler.schedule(obs)
        assert (action["stage_idx"], action["num_exec"]) == tuple(tr["actions"][k]), k
        obs, reward, terminated, truncated, info = env.step(action)
        assert reward == tr["reward"][k] and info["wall_time"] == tr["wall"][k]
        assert obs["dag_batch"].nodes.shape[0] == tr["N"][k + 1]
        assert len(obs["exec_supplies"]) == tr["Ja"][k + 1]
        k += 1
    assert k == len(tr["actions"])
    # executor.history as the renderer reads it (spark_sched_sim.py:411): [[t_release, job_id], ..., [None, job_id]]
    for e, ex in enumerate(env.executors):
        lo, hi = int(tr["hist_ptr"][e]), int(tr["hist_ptr"][e + 1])
        assert ex.history[0][1] == -1 and ex.history[-1][0] is None and len(ex.history) == hi - lo + 1
        assert [h[0] for h in ex.history[:-1]] == tr["hist_t"][lo:hi].tolist()
        assert [h[1] for h in ex.history[1:]] == tr["hist_job"][lo:hi].tolist()
    assert metrics.avg_job_duration(env) * 1e-3 == pytest.approx(
        np.mean(tr["job_t_completed"] - tr["job_t_arrival"]) * 1e-3, rel=1e-12)
    assert env.num_completed_jobs == len(tr["job_template"]) and env.all_jobs_complete
    # error conventions of the reference (spark_sched_sim.py:276-295)
    env.reset(seed=tr["seed"])
    with pytest.raises(ValueError):
        env.step({"stage_idx": 0, "num_exec": 0})
    with pytest.raises(ValueError):
        env.step({"stage_idx": 0})
    with pytest.raises(ValueError):
        env.step({"stage_idx": 0, "num_exec": cfg["num_executors"] + 1})
    with pytest.raises(ValueError):
        SparkSchedSimEnv(dict(cfg, job_arrival_cap=None), bank=bank).reset(seed=1)
    env.close()


def test_packed_observations_on_the_host(bank):
    """ssb_get_obs_host: the graphs of all envs packed back to back equal the per-env slabs (and the oracle's
    observation of the same episode)."""
    from oracle import OracleEnv
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    cfg = {"num_executors": 10, "job_arrival_cap": 7, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    B = 37
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank)
    seeds = np.arange(B, dtype=np.uint64) + 900
    env.reset_host(seeds)
    for it in range(40):
        a, n = env.fair_actions(True)
        hdr = env.step_host(a.cpu().numpy(), n.cpu().numpy()).copy()
        if it % 13 != 12:
            continue
        po = env.obs_host()
        o = po["offsets"]
        assert o.shape == (B + 1, 3) and (o[0] == 0).all()
        assert o[B].tolist() == [int(hdr["num_nodes"].sum()), int(hdr["num_edges"].sum()), int(hdr["num_active_jobs"].sum())]
        for b in range(B):
            ref = env.obs(b, hdr)
            assert np.array_equal(po["nodes"][o[b, 0]:o[b + 1, 0]], ref["nodes"])
            assert np.array_equal(po["edge_links"][o[b, 1]:o[b + 1, 1]], ref["edge_links"])
            assert np.array_equal(po["exec_supplies"][o[b, 2]:o[b + 1, 2]], ref["exec_supplies"])
            assert np.array_equal(po["dag_ptr"][o[b, 2] + b:o[b + 1, 2] + b + 1], ref["dag_ptr"])
    # env 5 against the oracle driven the same way
    orc = OracleEnv(bank, 10, 7, 2000.0, 1000.0, 4.0e-5)
    orc.reset_seed(int(seeds[5]))
    for it in range(39):  # the last packed copy was taken after call 39 (it == 38)
        orc.step(*orc.fair_action(True))
    oo, b = orc.obs(), 5
    assert np.array_equal(po["nodes"][o[b, 0]:o[b + 1, 0]], oo["nodes"])
    assert np.array_equal(po["edge_links"][o[b, 1]:o[b + 1, 1]], oo["edge_links"])


def test_executor_history_of_fused_rollouts_matches_oracle(bank):
    """ssb_get_history after whole episodes run by the fused rollout kernel == the oracle's add_history calls
    (Executor.add_history, executor.py:34-44), at 10 and at 50 executors; off by default (no rows, no cost)."""
    from oracle import OracleEnv
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    for E, J in ((10, 9), (50, 12)):
        cfg = {"num_executors": E, "job_arrival_cap": J, "job_arrival_rate": 4.0e-5,
               "moving_delay": 2000.0, "warmup_delay": 1000.0}
        B = 6
        env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank, history_capacity=4096)
        seeds = np.arange(B, dtype=np.uint64) + 300 + E
        env.reset_host(seeds)
        env.rollout_fair(100000, True, False)
        assert (env.hdr()["terminated"] == 1).all()
        for b in range(B):
            orc = OracleEnv(bank, E, J, 2000.0, 1000.0, 4.0e-5)
            orc.reset_seed(int(seeds[b]))
            term = False
            while not term:
                _, _, term = orc.step(*orc.fair_action(True))
            want, got = orc.history(), env.history(b)
            assert len(want["hist_t"]) > 0
            for k in ("hist_t", "hist_exec", "hist_job"):
                assert np.array_equal(want[k], got[k]), (E, b, k)
            hs = env.executor_histories(b)
            assert len(hs) == E and all(h[0][1] == -1 or len(h) == 1 for h in hs) and all(h[-1][0] is None for h in hs)
    plain = BatchedSparkSchedSimEnv(cfg, num_envs=2, bank=bank)
    plain.reset_host(np.array([1, 2], np.uint64))
    plain.rollout_fair(50, True, False)
    with pytest.raises(RuntimeError):
        plain.history(0)  # built without history_capacity: nothing was recorded


def test_config4_shape_batch_properties(bank):
    """BASELINE config 4's episode shape (200 jobs x 50 executors: two executor slots per lane) on a 1024-env batch:
    no env errors, every episode terminates, seed twins identical, deterministic, samples equal to the oracle."""
    from oracle import OracleEnv
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    B = 1024
    cfg = {"num_executors": 50, "job_arrival_cap": 200, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank)
    seeds = (4321 + np.arange(B) // 2).astype(np.uint64)
    walls = []
    for rep in range(2):
        env.reset_stats()
        env.reset_host(seeds)
        env.rollout_fair(1_000_000, True, auto_reset=False)
        hdr = env.hdr()
        assert (hdr["error"] == 0).all() and (hdr["terminated"] == 1).all()
        walls.append(hdr["wall_time"].copy())
    assert np.array_equal(walls[0], walls[1]) and np.array_equal(walls[0][0::2], walls[0][1::2])
    assert env.stats()["episodes"] == B
    for b in (0, 513, 1023):
        orc = OracleEnv(bank, 50, 200, 2000.0, 1000.0, 4.0e-5)
        orc.run_fair_episode(int(seeds[b]), True)
        assert np.array_equal(orc.job_times()[1], env.jobs(b)[1])
