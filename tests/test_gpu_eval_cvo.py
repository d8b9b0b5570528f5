"""accflow_b200/eval_cvo.py (the test_cvo.py-shaped caller of the hot path) on the real kernels vs the oracle."""
import pytest
import torch

from tests.golden import cases

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.mark.parametrize("acc,ofe", [("acc", "raft"), ("direct", "gma")])
def test_eval_driver_matches_oracle_metrics(tmp_path, capsys, acc, ofe):
    from accflow_b200 import eval_cvo
    from accflow_b200.data import preprocess
    from accflow_b200.dataset import fetch_valid_dataloader
    from oracle import flow_oracle as fo
    from oracle import ops
    kind = ("acc+" + ofe) if acc == "acc" else ofe
    sd = cases.weights(kind)
    ckpt = str(tmp_path / "ckpt.pth")
    torch.save({"module." + k: v for k, v in sd.items()}, ckpt)
    flag = "--acc_ckpt" if acc == "acc" else "--ofe_ckpt"
    res = eval_cvo.main(["-d", "clean", "-acc", acc, "-ofe", ofe, flag, ckpt, "--size", "128", "--clips", "5", "--batch", "2",
                         "--out-dir", str(tmp_path)])
    loader, _ = fetch_valid_dataloader(["fflows", "bflows"], "clean", batch=5, n_clips=5, size=128)
    data = preprocess(next(iter(loader)))
    imgs = data["imgs"]
    pred = fo.flow_estimator(sd, imgs[6], imgs[0], 12) if acc == "direct" else fo.accflow_forward(sd, imgs[:7], 12)[-1]
    occ, _ = ops.calc_occ_mask(data["bflows"][4], data["fflows"][4])
    want = torch.stack(ops.cal_epe(pred, data["bflows"][4], occ), 1)        # (5, 3): all, occ, vis
    assert float((res["per_clip"] - want).abs().max()) < 1e-4
    import re
    out = capsys.readouterr().out
    assert f"AVG EPE {acc}|{ofe}: " in out
    for text in (out, (tmp_path / "test_result_clean_E6.txt").read_text()):
        got = [float(x) for x in re.search(r"all:([\d.]+) vis:([\d.]+) occ:([\d.]+)", text).groups()]
        for a, b in zip(got, (want[:, 0].mean(), want[:, 2].mean(), want[:, 1].mean())):
            assert abs(a - float(b)) < 2e-4          # the script prints 4 decimals
