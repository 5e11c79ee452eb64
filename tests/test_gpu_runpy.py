"""The reference's UNMODIFIED codes/run.py (shipped to the GPU box as the git-ignored baseline/_ref/) running end to end
on the drop-in KGEModel with --cuda: train / valid / test on countries_S1 at BASELINE configs[0]'s shape, checkpoint +
optimizer resume, and --do_test -init on wn18rr.  Every metric line run.py logs is compared with the line the same
command logged on the reference itself (CPU) in the build container (tests/golden/runpy_logs.json, written by
tests/golden/make_golden.py --only runpy).  SURVEY section 4 item 4; run.py:227-242, 275-287, 303-364."""
import json
import os

import numpy as np
import pytest

import runpy_cases as RC

pytestmark = pytest.mark.gpu
REF = os.path.join(RC.ROOT, "baseline", "_ref")


def check(name, got, want):
    assert [r[:3] for r in got] == [r[:3] for r in want], (name, got, want)      # same lines, same order, same steps
    for (kind, metric, step, g), (_, _, _, w) in zip(got, want):
        if metric in ("MRR", "MR", "HITS@1", "HITS@3", "HITS@10"):
            # ranks may move across exact near-ties (libm / reduction order): same bound as test_wn18rr_real_dataset_ranks
            assert abs(g - w) <= 2e-4 * abs(w) + 2e-6, (name, kind, metric, step, g, w)
        else:
            assert abs(g - w) <= 1e-5 * abs(w) + 1.5e-6, (name, kind, metric, step, g, w)   # '%f' prints 6 decimals


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    if not os.path.exists(os.path.join(REF, "codes", "run.py")):
        pytest.skip("baseline/_ref (the staged reference, written by __graft_entry__.build()) is absent")
    work = str(tmp_path_factory.mktemp("runpy"))
    RC.prepare_wn18rr(os.path.join(REF, "data"), work)
    return work


def test_run_py_countries_train_valid_test_and_resume(workdir):
    golden = json.load(open(RC.GOLDEN_JSON))["logs"]
    cmds = RC.commands(os.path.join(REF, "data"), workdir)
    run_py = os.path.join(REF, "codes", "run.py")
    out = RC.run_case(run_py, cmds["countries_train"], reference=False, cuda=True)
    assert "Start Training" in out
    save = os.path.join(workdir, "countries")
    for name in ("config.json", "checkpoint", "entity_embedding.npy", "relation_embedding.npy", "train.log"):
        assert os.path.exists(os.path.join(save, name)), name
    assert np.load(os.path.join(save, "entity_embedding.npy")).shape == (271, 1000)
    check("countries_train", RC.parse_log(os.path.join(save, "train.log")), golden["countries_train"])
    # resume: run.py:275-287 loads model + optimizer state (LR already decayed once, Adam moments of the drop-in's run)
    RC.run_case(run_py, cmds["countries_resume"], reference=False, cuda=True)
    check("countries_resume", RC.parse_log(os.path.join(workdir, "countries_resumed", "train.log")), golden["countries_resume"])


def test_run_py_wn18rr_test_from_checkpoint(workdir):
    golden = json.load(open(RC.GOLDEN_JSON))["logs"]
    cmds = RC.commands(os.path.join(REF, "data"), workdir)
    RC.run_case(os.path.join(REF, "codes", "run.py"), cmds["wn18rr_test"], reference=False, cuda=True)
    check("wn18rr_test", RC.parse_log(os.path.join(workdir, "wn18rr_ckpt", "test.log")), golden["wn18rr_test"])
