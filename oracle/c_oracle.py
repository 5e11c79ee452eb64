"""ctypes loader for oracle/kge_oracle.c  --  TEST INFRASTRUCTURE (see the header of kge_oracle.c)."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libkge_oracle.so")
MODEL_IDS = {"TransE": 0, "DistMult": 1, "ComplEx": 2, "RotatE": 3, "pRotatE": 4}
MODE_IDS = {"single": 0, "head-batch": 1, "tail-batch": 2}
_lib = None


def _cpu_has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            return " fma " in f.read().replace("\n", " ")
    except OSError:
        return False


def build(force=False):
    src = os.path.join(HERE, "kge_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "libkge_oracle.so"], check=True, capture_output=True)
    return LIB


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = LIB
    if not os.path.exists(path):
        build()
    if not _cpu_has_fma():          # the in-tree build uses -mfma; rebuild portable code for an older host
        path = os.path.join(tempfile.mkdtemp(prefix="kge_oracle_"), "libkge_oracle.so")
        subprocess.run(["gcc", "-O2", "-std=c11", "-fPIC", "-fopenmp", "-ffp-contract=off", "-shared", "-o", path,
                        os.path.join(HERE, "kge_oracle.c"), "-lm"], check=True)
    _lib = ctypes.CDLL(path)
    _lib.ko_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def num_threads():
    return load().ko_num_threads()


def set_num_threads(n):
    load().ko_set_num_threads(int(n))


def sincos(x):
    x = _f32(x)
    s, c = np.empty_like(x), np.empty_like(x)
    load().ko_sincos(_p(x), ctypes.c_int64(x.size), _p(s), _p(c))
    return s, c


def _modulus(state):
    return float(np.asarray(state["modulus"]).reshape(-1)[0]) if "modulus" in state else 1.0


def query_vectors(model, state, queries, mode, rho):
    E, R = _f32(state["entity_embedding"]), _f32(state["relation_embedding"])
    q = _i64(queries).reshape(-1, 3)
    out = np.empty((q.shape[0], E.shape[1]), dtype=np.float32)
    load().ko_query_vectors(MODEL_IDS[model], int(mode == "head-batch"), _p(E), _p(R), E.shape[1], R.shape[1],
                            ctypes.c_float(rho), _p(q), ctypes.c_int64(q.shape[0]), _p(out))
    return out


def forward(model, state, sample, mode, gamma, rho):
    E, R = _f32(state["entity_embedding"]), _f32(state["relation_embedding"])
    if mode == "single":
        pos, neg, N = _i64(sample), None, 1
    else:
        pos, neg = _i64(sample[0]), _i64(sample[1])
        N = neg.shape[1]
    out = np.empty((pos.shape[0], N), dtype=np.float32)
    load().ko_forward(MODEL_IDS[model], MODE_IDS[mode], _p(E), _p(R), E.shape[1], R.shape[1], ctypes.c_float(gamma),
                      ctypes.c_float(rho), ctypes.c_float(_modulus(state)), _p(pos), _p(neg),
                      ctypes.c_int64(pos.shape[0]), ctypes.c_int64(N), _p(out))
    return out


def eval_scores(model, state, queries, mode, gamma, rho, csr_offsets, csr_entities):
    """[Q, nentity] scores + filter_bias exactly as model.py:392-393 sees them."""
    E, R = _f32(state["entity_embedding"]), _f32(state["relation_embedding"])
    q = _i64(queries).reshape(-1, 3)
    off = _i64(csr_offsets)
    ent = np.ascontiguousarray(csr_entities, dtype=np.int32)
    out = np.empty((q.shape[0], E.shape[0]), dtype=np.float32)
    load().ko_eval_scores(MODEL_IDS[model], int(mode == "head-batch"), _p(E), _p(R), ctypes.c_int64(E.shape[0]),
                          E.shape[1], R.shape[1], ctypes.c_float(gamma), ctypes.c_float(rho),
                          ctypes.c_float(_modulus(state)), _p(q), ctypes.c_int64(q.shape[0]), _p(off), _p(ent), _p(out))
    return out


def ranks_from_scores(scores, queries, mode):
    s = _f32(scores)
    q = _i64(queries).reshape(-1, 3)
    out = np.empty(q.shape[0], dtype=np.int64)
    load().ko_ranks_from_scores(_p(s), ctypes.c_int64(s.shape[0]), ctypes.c_int64(s.shape[1]), _p(q),
                                int(mode == "head-batch"), _p(out))
    return out


class TrainState:
    """Tables + Adam moments for ko_train_step (mirrors oracle.kge_oracle.TrainState)."""

    def __init__(self, model_name, state, gamma, hidden_dim):
        from . import kge_oracle as O
        self.model_name, self.gamma, self.hidden_dim = model_name, float(gamma), int(hidden_dim)
        self.rho = O.embedding_range(gamma, hidden_dim)
        self.state = {k: np.array(v, dtype=np.float32) for k, v in state.items()}
        self.reset_optimizer()

    def reset_optimizer(self):
        self.step = 0
        self.m = {k: np.zeros_like(v) for k, v in self.state.items()}
        self.v = {k: np.zeros_like(v) for k, v in self.state.items()}
        self.grads = {k: np.zeros_like(v) for k, v in self.state.items()}


def train_step(ts, batch, *, lr, adversarial, alpha=1.0, uni_weight=False, regularization=0.0, return_grads=False):
    positive, negative, weight, mode = batch
    pos, neg = _i64(positive), _i64(negative)
    w = None if uni_weight else _f32(weight)
    E, R = ts.state["entity_embedding"], ts.state["relation_embedding"]
    M = ts.state.get("modulus")
    out = np.zeros(4, dtype=np.float32)
    ts.step += 1
    k = ("entity_embedding", "relation_embedding", "modulus")
    load().ko_train_step(
        MODEL_IDS[ts.model_name], MODE_IDS[mode], _p(E), _p(R), _p(M), ctypes.c_int64(E.shape[0]),
        ctypes.c_int64(R.shape[0]), E.shape[1], R.shape[1], ctypes.c_float(ts.gamma), ctypes.c_float(ts.rho), _p(pos),
        _p(neg), _p(w), ctypes.c_int64(pos.shape[0]), ctypes.c_int64(neg.shape[1]), int(bool(adversarial)),
        ctypes.c_float(alpha), ctypes.c_double(regularization), ctypes.c_double(lr), int(ts.step),
        _p(ts.m[k[0]]), _p(ts.v[k[0]]), _p(ts.m[k[1]]), _p(ts.v[k[1]]),
        _p(ts.m.get(k[2])), _p(ts.v.get(k[2])), _p(ts.grads[k[0]]), _p(ts.grads[k[1]]), _p(ts.grads.get(k[2])), _p(out))
    log = {}
    if regularization != 0.0:
        log["regularization"] = float(out[3])
    log.update(positive_sample_loss=float(out[0]), negative_sample_loss=float(out[1]), loss=float(out[2]))
    return (log, {n: g.copy() for n, g in ts.grads.items()}) if return_grads else log
