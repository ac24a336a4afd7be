"""ctypes front-end of the CPU oracle (oracle/sim_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline / `--impl reference` legs import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import os.path as osp
import subprocess

import numpy as np

HERE = osp.dirname(osp.abspath(__file__))
LIB_PATH = osp.join(HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    src = [osp.join(HERE, "sim_oracle.c"), osp.join(HERE, "sim_oracle.h")]
    if (
        force
        or not osp.exists(LIB_PATH)
        or any(osp.getmtime(s) > osp.getmtime(LIB_PATH) for s in src)
    ):
        subprocess.check_call(["make", "-C", HERE, "-B", "_build/liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return LIB_PATH


class _Cfg(C.Structure):
    _fields_ = [
        ("num_executors", C.c_int32),
        ("job_arrival_cap", C.c_int32),
        ("moving_delay", C.c_double),
        ("warmup_delay", C.c_double),
        ("job_arrival_rate", C.c_double),
        ("beta", C.c_double),
    ]


class _Bank(C.Structure):
    _fields_ = [
        ("num_templates", C.c_int32),
        ("num_stages", C.c_void_p),
        ("stage_base", C.c_void_p),
        ("edge_base", C.c_void_p),
        ("edges", C.c_void_p),
        ("num_tasks", C.c_void_p),
        ("rough_duration", C.c_void_p),
        ("present", C.c_void_p),
        ("dur_off", C.c_void_p),
        ("dur_cnt", C.c_void_p),
        ("dur_values", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(_Cfg), C.POINTER(_Bank)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_reset_trace.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_uint64]
        L.orc_reset_seed.argtypes = [C.c_void_p, C.c_uint64, C.c_double]
        L.orc_step.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double),
                               C.POINTER(C.c_int32)]
        L.orc_obs_sizes.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_obs_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.orc_fair_action.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.orc_wall_time.restype = C.c_double
        L.orc_wall_time.argtypes = [C.c_void_p]
        L.orc_num_jobs.argtypes = [C.c_void_p]
        L.orc_error.argtypes = [C.c_void_p]
        L.orc_num_launches.restype = C.c_int64
        L.orc_num_launches.argtypes = [C.c_void_p]
        L.orc_job_times.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.orc_log_enable.argtypes = [C.c_void_p, C.c_int32]
        L.orc_log_size.restype = C.c_int64
        L.orc_log_size.argtypes = [C.c_void_p]
        L.orc_log_copy.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 7
        L.orc_history_size.restype = C.c_int64
        L.orc_history_size.argtypes = [C.c_void_p]
        L.orc_history_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.orc_run_fair_episode.restype = C.c_int64
        L.orc_run_fair_episode.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.POINTER(C.c_int64)]
        L.orc_philox4x32_10.argtypes = [C.c_void_p] * 3
        L.orc_neglog_u32.restype = C.c_double
        L.orc_neglog_u32.argtypes = [C.c_uint32]
        L.orc_pyset_new.restype = C.c_void_p
        L.orc_pyset_free.argtypes = [C.c_void_p]
        L.orc_pyset_add.argtypes = [C.c_void_p, C.c_int32]
        L.orc_pyset_remove.argtypes = [C.c_void_p, C.c_int32]
        L.orc_pyset_pop.argtypes = [C.c_void_p]
        L.orc_pyset_copy.restype = C.c_void_p
        L.orc_pyset_copy.argtypes = [C.c_void_p]
        L.orc_pyset_len.argtypes = [C.c_void_p]
        L.orc_pyset_list.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleEnv:
    """Single-environment CPU oracle with the reference's reset/step contract."""

    def __init__(self, bank, num_executors, job_arrival_cap, moving_delay, warmup_delay,
                 job_arrival_rate, beta=0.0, log=False):
        self.L = lib()
        self.bank = bank
        self._keep = [np.ascontiguousarray(x) for x in (
            bank.num_stages.astype(np.int32), bank.stage_base.astype(np.int32),
            bank.edge_base.astype(np.int32), bank.edges.astype(np.int32),
            bank.num_tasks.astype(np.int32), bank.rough_duration.astype(np.float64),
            bank.present.astype(np.uint8), bank.dur_off.astype(np.uint32),
            bank.dur_cnt.astype(np.uint32), bank.dur_values.astype(np.float64))]
        b = _Bank(bank.num_templates, *[_p(x) for x in self._keep])
        cfg = _Cfg(int(num_executors), int(job_arrival_cap or 0), float(moving_delay),
                   float(warmup_delay), float(job_arrival_rate), float(beta))
        self.h = self.L.orc_create(C.byref(cfg), C.byref(b))
        if not self.h:
            raise RuntimeError("orc_create failed")
        self.E = int(num_executors)
        if log:
            self.L.orc_log_enable(self.h, 1)

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def reset_trace(self, t_arrival, template, tape=None, seed=0):
        ta = np.ascontiguousarray(t_arrival, np.float64)
        tm = np.ascontiguousarray(template, np.int32)
        if tape is not None:
            tp = np.ascontiguousarray(tape, np.float64)
            rc = self.L.orc_reset_trace(self.h, len(ta), _p(ta), _p(tm), _p(tp), len(tp), seed)
        else:
            rc = self.L.orc_reset_trace(self.h, len(ta), _p(ta), _p(tm), None, 0, seed)
        if rc:
            raise RuntimeError(f"oracle reset error {rc}")
        return self.obs()

    def reset_seed(self, seed, time_limit=np.inf):
        rc = self.L.orc_reset_seed(self.h, int(seed), float(time_limit))
        if rc:
            raise RuntimeError(f"oracle reset error {rc}")
        return self.obs()

    def step(self, stage_idx, num_exec):
        r = C.c_double()
        t = C.c_int32()
        rc = self.L.orc_step(self.h, int(stage_idx), int(num_exec), C.byref(r), C.byref(t))
        return rc, r.value, bool(t.value)

    def obs(self):
        sc = np.zeros(5, np.int32)
        self.L.orc_obs_sizes(self.h, _p(sc))
        N, M, Ja, ncommit, src = (int(x) for x in sc)
        nodes = np.zeros((N, 3), np.float32)
        edges = np.zeros((M, 2), np.int32)
        dag_ptr = np.zeros(Ja + 1, np.int32)
        sup = np.zeros(Ja, np.int32)
        self.L.orc_obs_copy(self.h, _p(nodes), _p(edges), _p(dag_ptr), _p(sup))
        return {"nodes": nodes, "edge_links": edges, "dag_ptr": dag_ptr, "exec_supplies": sup,
                "num_committable_execs": ncommit, "source_job_idx": src}

    def fair_action(self, dynamic_partition=True):
        a, n = C.c_int32(), C.c_int32()
        self.L.orc_fair_action(self.h, int(dynamic_partition), C.byref(a), C.byref(n))
        return a.value, n.value

    @property
    def wall_time(self):
        return self.L.orc_wall_time(self.h)

    @property
    def num_launches(self):
        return self.L.orc_num_launches(self.h)

    def job_times(self):
        n = self.L.orc_num_jobs(self.h)
        ta, tc, tm = np.zeros(n), np.zeros(n), np.zeros(n, np.int32)
        self.L.orc_job_times(self.h, _p(ta), _p(tc), _p(tm))
        return ta, tc, tm

    def log(self, lo=0, hi=None):
        n = self.L.orc_log_size(self.h)
        hi = n if hi is None else hi
        k = hi - lo
        out = {"ev_t": np.zeros(k), "ev_type": np.zeros(k, np.uint8), "ev_job": np.zeros(k, np.int16),
               "ev_stage": np.zeros(k, np.int16), "ev_task": np.zeros(k, np.int32),
               "ev_exec": np.zeros(k, np.int16), "ev_tacc": np.zeros(k)}
        self.L.orc_log_copy(self.h, lo, hi, *[_p(out[x]) for x in (
            "ev_t", "ev_type", "ev_job", "ev_stage", "ev_task", "ev_exec", "ev_tacc")])
        return out

    def history(self):
        """Every Executor.add_history call of the episode in call order: (t, executor id, job id or -1)."""
        n = self.L.orc_history_size(self.h)
        t, ex, jb = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)
        if n:
            self.L.orc_history_copy(self.h, _p(t), _p(ex), _p(jb))
        return {"hist_t": t, "hist_exec": ex, "hist_job": jb}

    def log_size(self):
        return self.L.orc_log_size(self.h)

    def run_fair_episode(self, seed, dynamic_partition=True):
        ev = C.c_int64()
        n = self.L.orc_run_fair_episode(self.h, int(seed), int(dynamic_partition), C.byref(ev))
        if n < 0:
            raise RuntimeError(f"oracle episode error {-n}")
        return int(n), int(ev.value)
