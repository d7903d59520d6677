"""Device-side generation of the synthetic workloads (cilqr_b200_synth_generate) against the numpy generator
(scenario.generate_host): the same arrays bit for bit for every BASELINE config, any slice of the instance-id
range, and therefore the same solves."""
import numpy as np
import pytest

import cilqr_b200 as cb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg,N", [("C1", 50), ("C2", 100), ("C3", 50), ("C4", 200)])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_device_generated_equals_host_generated(cfg, N, dtype):
    B, first = 4096, 123456
    spec = cb.synth_spec(cfg, N)
    host = cb.generate_host(spec, B, first_id=first)
    with cb.BatchSolver(spec.templates, B, N, spec.max_obs, dtype) as s:
        s.generate(spec, B, first_id=first, keep_yaw=True)
        x0, rv, bd, tm, no, ob = s.synth_download(B)
    cast = (lambda a: a.astype(np.float32).astype(np.float64)) if dtype == "f32" else (lambda a: a)
    assert np.array_equal(tm, host.tmpl) and np.array_equal(no, host.n_obs)
    assert np.array_equal(x0, cast(host.x0))
    assert np.array_equal(rv, cast(host.ref_velo))
    assert np.array_equal(bd, cast(host.borders))
    assert np.array_equal(ob[..., :3], cast(host.obs))      # x, y and the raw yaw
    assert np.all(ob[..., 3] == 0)
    assert np.abs(host.obs[..., 2]).max() > 3.0 or cfg in ("C1", "C2")  # oncoming traffic (yaw + pi) is exercised


@pytest.mark.parametrize("cfg,N,dtype", [("C1", 50, "f64"), ("C3", 50, "f64"), ("C2", 100, "f64"), ("C4", 200, "f32")])
def test_solving_a_device_generated_batch_equals_solving_the_uploaded_one(cfg, N, dtype):
    B, first = 1024, 777
    spec = cb.synth_spec(cfg, N)
    host = cb.generate_host(spec, B, first_id=first)
    with cb.BatchSolver(spec.templates, B, N, spec.max_obs, dtype) as s:
        a = s.solve(host)
        s.generate(spec, B, first_id=first)
        s.solve_resident(B)
        b = s.download(B)
    for f in ("u", "x", "J", "K", "d", "iters", "status", "exit_reason", "step_cost"):
        assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True), f
