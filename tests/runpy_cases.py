"""The `codes/run.py` command lines shared by tests/golden/make_golden.py (run on the UNMODIFIED reference, CPU) and
tests/test_gpu_runpy.py (same commands with --cuda on the drop-in), plus the log parser.

Cases (run.py:227-242, 275-287, 303-364):
  countries_train    run.py --do_train --do_valid --do_test --countries on countries_S1 (BASELINE configs[0] shape), 20 steps
  countries_resume   run.py --do_train --do_test -init <that run's directory>: checkpoint + optimizer state round trip
  wn18rr_test        run.py --do_test -init <checkpoint written from seeded tables> on wn18rr (first 200 test triples)
"""
import json
import os
import re
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAUNCHER = os.path.join(ROOT, "knowledgegraphembedding_b200", "dropin", "run_with_dropin.py")
GOLDEN_JSON = os.path.join(ROOT, "tests", "golden", "runpy_logs.json")

COUNTRIES = ["--data_path", "{data}/countries_S1", "--model", "RotatE", "-n", "64", "-b", "512", "-d", "500", "-g", "0.1",
             "-a", "1.0", "-adv", "-lr", "0.000002", "-de", "--countries", "-cpu", "2", "--test_batch_size", "8",
             "--log_steps", "5", "--valid_steps", "10", "--save_checkpoint_steps", "10"]


def commands(data, work):
    """name -> run.py argument list (without --cuda)."""
    c = [a.format(data=data) for a in COUNTRIES]
    return {
        "countries_train": ["--do_train", "--do_valid", "--do_test", "--max_steps", "20", "-save", f"{work}/countries"] + c,
        "countries_resume": ["--do_train", "--do_test", "--max_steps", "30", "-init", f"{work}/countries",
                             "-save", f"{work}/countries_resumed"] + c,
        "wn18rr_test": ["--do_test", "-init", f"{work}/wn18rr_ckpt", "--data_path", f"{work}/wn18rr_small", "-cpu", "2"],
    }


def prepare_wn18rr(data, work, nq=200, d=32, gamma=6.0, seed=11, scale=6.0):
    """A data directory with the first `nq` test triples of wn18rr and a run.py checkpoint directory whose tables come
    from the portable numpy initialiser (no training needed: the point is run.py's load + test_step + logging path)."""
    import torch
    sys.path.insert(0, ROOT)
    from oracle import kge_oracle as O
    src, dst = os.path.join(data, "wn18rr"), os.path.join(work, "wn18rr_small")
    os.makedirs(dst, exist_ok=True)
    for name in ("entities.dict", "relations.dict", "train.txt", "valid.txt"):
        shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
    with open(os.path.join(src, "test.txt")) as f, open(os.path.join(dst, "test.txt"), "w") as g:
        g.writelines(f.readlines()[:nq])
    nentity = sum(1 for _ in open(os.path.join(src, "entities.dict")))
    nrel = sum(1 for _ in open(os.path.join(src, "relations.dict")))
    st = O.init_tables("RotatE", nentity, nrel, d, gamma, True, False, seed=seed)
    ckpt = os.path.join(work, "wn18rr_ckpt")
    os.makedirs(ckpt, exist_ok=True)
    with open(os.path.join(ckpt, "config.json"), "w") as f:
        json.dump({"countries": False, "data_path": dst, "model": "RotatE", "double_entity_embedding": True,
                   "double_relation_embedding": False, "hidden_dim": d, "test_batch_size": 16}, f)
    state = {"gamma": torch.tensor([gamma]), "embedding_range": torch.tensor([(gamma + 2.0) / d]),
             "entity_embedding": torch.from_numpy((st["entity_embedding"] * scale).astype(np.float32)),
             "relation_embedding": torch.from_numpy(st["relation_embedding"])}
    torch.save({"step": 7, "current_learning_rate": 1e-4, "warm_up_steps": 100, "model_state_dict": state,
                "optimizer_state_dict": {}}, os.path.join(ckpt, "checkpoint"))


def run_case(run_py, argv, reference, cuda, seed=0, timeout=1500):
    env = dict(os.environ, KGE_RUN_SEED=str(seed))
    env.pop("KGE_RUN_REFERENCE", None)
    if reference:
        env["KGE_RUN_REFERENCE"] = "1"
    cmd = [sys.executable, LAUNCHER, run_py] + (["--cuda"] if cuda else []) + list(argv)
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    if res.returncode != 0:
        raise RuntimeError("run.py failed (%s):\n%s" % (" ".join(cmd), (res.stdout + res.stderr)[-4000:]))
    return res.stdout + res.stderr


METRIC_LINE = re.compile(r"(Training average|Valid|Test) (\S+) at step (\d+): (-?[0-9.]+(?:e-?\d+)?|nan|inf)")


def parse_log(path):
    """[(kind, metric, step, value)] of run.py's log_metrics lines (run.py:164-169), in order."""
    out = []
    with open(path) as f:
        for line in f:
            m = METRIC_LINE.search(line)
            if m:
                out.append([m.group(1), m.group(2), int(m.group(3)), float(m.group(4))])
    return out
