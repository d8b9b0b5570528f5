"""The reference's own evaluation script (test_cvo.py, UNMODIFIED, executed with runpy) on this package:
``networks`` -> ``accflow_b200.networks``, ``data.dataset`` -> ``accflow_b200.dataset`` (synthetic CVO-shaped
records), checkpoints written from the seeded weights through the ``module.`` prefix its DataParallel wrapper
expects (test_cvo.py:11-29, 132-166).

Runs where the reference checkout exists, i.e. in the build container — which has no GPU, while the product has no
CPU path.  So exactly three things are stood in, and only here: ``.cuda()`` is the identity, the CUDA engine behind
``AccFlow.engine()`` / ``RAFT.engine()`` is replaced by an adaptor that calls the CPU oracle with the module's own
``state_dict()`` (tests may use the oracle), and ``networks.utils.backwarp`` (a CUDA op here) by the oracle's.
Everything the script touches on the host side is the product's: ``build_flow_estimator``, the module tree and its
state-dict contract, ``nn.DataParallel(...).load_state_dict``, the keyword call ``model(images=..., test_mode=False)``,
the positional ``model(imgN, img0)``, the record layout of the loader, output shapes/dtypes fed to the script's own
``calc_occ_mask`` / ``cal_epe``, and the result file.  The same driver logic on the real kernels is
``accflow_b200/eval_cvo.py`` (tests/test_gpu_eval_cvo.py)."""
import os
import runpy
import sys
import types

import pytest
import torch

from tests.golden import cases

REF_SCRIPT = "/root/reference/test_cvo.py"
pytestmark = pytest.mark.skipif(not os.path.isfile(REF_SCRIPT), reason="reference checkout not present")
torch.set_grad_enabled(False)

SIZE, CLIPS = 128, 3


class _OracleAccEngine:
    def __init__(self, module):
        self.m = module

    def forward(self, images, iters, graph=False, warm_start=False, warm_iters=None):
        from oracle import flow_oracle as fo
        return fo.accflow_forward(dict(self.m.state_dict()), list(images), iters, warm_start=warm_start, warm_iters=warm_iters)


class _OraclePairEngine:
    def __init__(self, module):
        self.m = module

    def forward(self, image1, image2, iters, flow_init, graph=False):
        from oracle import flow_oracle as fo
        return fo.flow_estimator(dict(self.m.state_dict()), image1, image2, iters, flow_init)


@pytest.fixture
def swapped(monkeypatch, tmp_path):
    import accflow_b200.dataset as ds
    import accflow_b200.networks as nets
    import accflow_b200.networks.AccFlow_ as accmod
    import accflow_b200.networks.utils as nutils
    from accflow_b200.networks._estimator import FlowEstimatorBase
    from oracle import ops
    monkeypatch.setitem(sys.modules, "networks", nets)
    monkeypatch.setitem(sys.modules, "networks.AccFlow_", accmod)
    monkeypatch.setitem(sys.modules, "networks.utils", nutils)
    datapkg = types.ModuleType("data")
    datapkg.dataset = ds
    monkeypatch.setitem(sys.modules, "data", datapkg)
    monkeypatch.setitem(sys.modules, "data.dataset", ds)
    # stand-ins for the missing GPU (see module docstring)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(accmod.AccFlow, "engine", lambda self, device=None: _OracleAccEngine(self))
    monkeypatch.setattr(FlowEstimatorBase, "engine", lambda self, device=None: _OraclePairEngine(self))
    monkeypatch.setattr(nutils, "backwarp", ops.backwarp)
    monkeypatch.setenv("ACCFLOW_CVO_CLIPS", str(CLIPS))
    monkeypatch.setenv("ACCFLOW_CVO_SIZE", str(SIZE))
    monkeypatch.chdir(tmp_path)
    return tmp_path


def _expected(kind, direct):
    from accflow_b200.data import preprocess
    from accflow_b200.dataset import fetch_valid_dataloader
    from oracle import flow_oracle as fo
    from oracle import ops
    sd = cases.weights(kind)
    loader, _ = fetch_valid_dataloader(["fflows", "bflows"], "clean", batch=10, n_clips=CLIPS, size=SIZE)
    (batch,) = list(loader)
    data = preprocess(batch)
    imgs = data["imgs"]
    if direct:
        pred = fo.flow_estimator(sd, imgs[6], imgs[0], 12)
    else:
        pred = fo.accflow_forward(sd, imgs[:7], 12)[-1]
    occ, _ = ops.calc_occ_mask(data["bflows"][4], data["fflows"][4])
    e_all, e_occ, e_vis = ops.cal_epe(pred, data["bflows"][4], occ)
    return "all:%.4f vis:%.4f occ:%.4f" % (e_all.mean(), e_vis.mean(), e_occ.mean())


@pytest.mark.parametrize("acc,ofe", [("acc", "raft"), ("direct", "gma")])
def test_reference_test_cvo_unmodified(swapped, monkeypatch, capsys, acc, ofe):
    kind = ("acc+" + ofe) if acc == "acc" else ofe
    ckpt = str(swapped / "ckpt.pth")
    torch.save({"module." + k: v for k, v in cases.weights(kind).items()}, ckpt)      # as train_acc.py:109 writes it
    flag = "--acc_ckpt" if acc == "acc" else "--ofe_ckpt"
    monkeypatch.setattr(sys, "argv", ["test_cvo.py", "-d", "clean", "-acc", acc, "-ofe", ofe, flag, ckpt])
    runpy.run_path(REF_SCRIPT, run_name="__main__")
    out = capsys.readouterr().out
    want = _expected(kind, direct=(acc == "direct"))
    assert f"AVG EPE {acc}|{ofe}: " in out
    assert want in out, (want, out[-300:])
    assert want in (swapped / "test_result_clean_E6.txt").read_text()
