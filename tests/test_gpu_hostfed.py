"""Host-fed pipelined stepper (trainer.HostFedStepper): the DataLoader-style loop of src/train.py:158-192 with host
batches.  It must produce exactly what ChainEngine.train_step produces on the same batches (the kernels are
deterministic, so losses and parameters are compared bit for bit), including a short last batch and a row mask."""
import pytest
import torch

from oracle import golden_util as G
from oracle.cases import case_setup

pytestmark = pytest.mark.gpu


def _engine(inr, name, bs):
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup(name)
    plan = inr.Plan(model_kind, net, enc_cfg)
    eng = inr.ChainEngine(plan, max_batch=bs, lr=G.LR)
    eng.load_tensors(list(sd.values()))
    eng.set_encoder(encB)
    return eng, loss_kind, opts, coords, gt


@pytest.mark.parametrize("name,masked", [("siren_l2", False), ("siren_l2", True), ("wire_hdr", True)])
def test_hostfed_equals_eager(name, masked):
    import mri_implicit_neural_representations_b200 as inr
    from mri_implicit_neural_representations_b200.trainer import HostFedStepper
    bs = 300
    eng_a, loss_kind, opts, coords, gt = _engine(inr, name, bs)
    eng_b, *_ = _engine(inr, name, bs)
    n = coords.shape[0]
    batches = [(i, min(bs, n - i)) for i in range(0, n, bs)] * 3        # three passes: eager, eager + capture, replay
    assert batches[-1][1] != bs or n % bs == 0
    mask = (torch.arange(n) % 3 != 0).to(torch.uint8) if masked else None
    # A: plain eager steps, loss read after every step
    ref = []
    for i, b in batches:
        eng_a.train_step(loss_kind, coords[i:i + b].cuda(), gt[i:i + b].cuda(), b,
                         mask=None if mask is None else mask[i:i + b].cuda(), loss_opts=opts)
        ref.append(float(eng_a.loss_out))
    # B: host batches through the pipelined stepper, loss read one step late
    st = HostFedStepper(eng_b, loss_kind, bs, masked=masked, loss_opts=opts, depth=2)
    got, prev = [], None
    for i, b in batches:
        t = st.submit(coords[i:i + b], gt[i:i + b], None if mask is None else mask[i:i + b])
        if prev is not None:
            got.append(st.loss(prev))
        prev = t
    got.append(st.loss(prev))
    assert len(st._graphs) >= 2, "the CUDA-graph path did not engage"
    assert got == ref
    torch.cuda.synchronize()
    assert torch.equal(eng_a.params, eng_b.params)
    assert st.h2d_bytes >= bs * (20 + (1 if masked else 0))


def test_hostfed_rejects_bad_input():
    import mri_implicit_neural_representations_b200 as inr
    from mri_implicit_neural_representations_b200.trainer import HostFedStepper
    eng, loss_kind, opts, coords, gt = _engine(inr, "siren_l2", 128)
    with pytest.raises(inr.InrError):
        HostFedStepper(eng, "CenterLoss", 128)
    st = HostFedStepper(eng, "L2", 128, masked=True)
    with pytest.raises(inr.InrError):
        st.submit(coords[:256], gt[:256], torch.ones(256, dtype=torch.uint8))
    with pytest.raises(inr.InrError):
        st.submit(coords[:128], gt[:128])
