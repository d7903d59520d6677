"""Batches beyond a million instances on one GPU (the verdict kernel's look-back runs over > 8192 chunks, one CTA
each; round 1 hung here above 606 208 instances before that fix): every instance exits, and a slice solved on its
own returns the same bits.  The batch is generated on the device (no host arrays of that size)."""
import numpy as np
import pytest

import cilqr_b200 as cb

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(600)
@pytest.mark.parametrize("B,dtype", [(1 << 20, "f64"), (3 << 19, "f32")])
def test_million_instances_on_one_gpu(B, dtype):
    spec = cb.synth_spec("C1", 50)
    lo = B // 2 + 12345
    with cb.BatchSolver(spec.templates, B, 50, spec.max_obs, dtype) as s:
        s.generate(spec, B)
        s.solve_resident(B)
        out = s.download(B, want_gains=False)
        c = s.counters()  # (the exit histogram is filled by the download)
        s.generate(spec, 256, first_id=lo)
        s.solve_resident(256)
        sub = s.download(256, want_gains=False)
    assert sum(c["exits"].values()) == B
    assert out.iters.min() >= 1 and out.iters.max() <= 100
    assert c["total_trials"] >= c["total_iters"] - int((out.status == 2).sum()) * 100
    for f in ("u", "x", "J", "iters", "status", "exit_reason"):
        assert np.array_equal(getattr(out, f)[lo:lo + 256], getattr(sub, f), equal_nan=True), f


def test_batch_limit_is_enforced():
    spec = cb.synth_spec("C1", 50)
    with pytest.raises(cb.CilqrError) as e:
        cb.BatchSolver(spec.templates, (1 << 23), 50, spec.max_obs, "f32")
    assert e.value.code == -1 and "max_batch" in str(e.value)
