"""CPU checks of two numerical claims the CUDA path relies on (no GPU needed):

* m_exp(double) in csrc/cilqr_model.cuh — the branch-free exp of the barrier terms — is within 1 ulp of glibc's exp
  and keeps inf / 0 / NaN: the function body is lifted out of the CUDA header as text and compiled with gcc
  (fma() from libm for __fma_rn), so this tests the shipped source, not a copy.
* the sufficient condition tried before the exact LLT test in riccati_step never says "positive definite" where
  the exact sequence (Eigen::LLT's arithmetic) says it is not.
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODEL = os.path.join(ROOT, "toy-example-of-ilqr_b200", "csrc", "cilqr_model.cuh")
KERNELS = os.path.join(ROOT, "toy-example-of-ilqr_b200", "csrc", "cilqr_kernels.cuh")


def _function_body(src, signature):
    i = src.index(signature)
    j = src.index("{", i)
    depth, k = 0, j
    while True:
        depth += src[k] == "{"
        depth -= src[k] == "}"
        if depth == 0:
            return src[i:k + 1]
        k += 1


def test_branch_free_exp_against_libm(tmp_path):
    body = _function_body(open(MODEL).read(), "__device__ __forceinline__ double m_exp(double x)")
    body = body.replace("__device__ __forceinline__ ", "static ").replace("__fma_rn", "fma")
    prog = r"""
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static int __double2loint(double d) { uint64_t b; memcpy(&b, &d, 8); return (int)(uint32_t)b; }
static double __hiloint2double(int hi, int lo) { uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &b, 8); return d; }
%s
static uint64_t rng = 88172645463325252ull;
static double uni(double a, double b) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return a + (b - a) * (double)(rng >> 11) / 9007199254740992.0; }
int main(void) {
    long worst = 0;
    for (int i = 0; i < 3000000; ++i) {
        double x = (i %% 3 == 0) ? uni(-700, 700) : (i %% 3 == 1) ? uni(-40, 40) : uni(-2, 2);
        double a = m_exp(x), b = exp(x);
        int64_t ia, ib; memcpy(&ia, &a, 8); memcpy(&ib, &b, 8);
        long d = labs((long)(ia - ib));
        if (d > worst) worst = d;
    }
    double sp[] = {0.0, -0.0, 709.7, 709.79, 710.0, 745.0, -745.0, -746.0, -800.0, 1e308, -1e308, INFINITY, -INFINITY};
    int bad = 0;
    for (unsigned i = 0; i < sizeof sp / sizeof *sp; ++i) {
        double a = m_exp(sp[i]), b = exp(sp[i]);
        if (!(a == b)) bad++;
    }
    if (!isnan(m_exp(NAN))) bad++;
    printf("%%ld %%d\n", worst, bad);
    return 0;
}
""" % body
    c = tmp_path / "exp_check.c"
    c.write_text(prog)
    exe = tmp_path / "exp_check"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(c), "-lm"], check=True)
    worst, bad = map(int, subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split())
    assert worst <= 1, "max distance to glibc exp: %d ulp" % worst
    assert bad == 0, "special values differ from libm"


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_pd_filter_never_overrules_the_exact_llt_test(dtype):
    src = open(KERNELS).read()
    # the margin the kernel uses (so that the test follows the source)
    m = re.search(r"surely_pd = Quu00 > T\(0\) && \(Quu00 \* Quu11 - a10sq\) > T\((\d+)\) \* kEps<T>\(\) \* a10sq", src)
    assert m, "riccati_step's filter not found"
    margin = dtype(int(m.group(1))) * np.finfo(dtype).eps
    rng = np.random.default_rng(5)
    n = 2_000_000
    a00 = (10.0 ** rng.uniform(-6, 6, n)).astype(dtype)
    a10 = ((10.0 ** rng.uniform(-6, 6, n)) * rng.choice([-1.0, 1.0], n)).astype(dtype)
    # a11 on and around the singular boundary a10^2 / a00, from far inside to far outside
    rel = np.concatenate([np.zeros(n // 4), rng.normal(0, 1, n // 4) * 1e-15, rng.normal(0, 1, n // 4) * 1e-6,
                          rng.uniform(-1, 1, n - 3 * (n // 4))])
    a11 = (a10.astype(np.float64) ** 2 / a00.astype(np.float64) * (1 + rel)).astype(dtype)
    with np.errstate(all="ignore"):
        l10 = a10 / np.sqrt(a00)
        exact_fails = (a00 <= 0) | (a11 - l10 * l10 <= 0)
        a10sq = a10 * a10
        # both ways the device may round a00 * a11 - a10^2: separate multiply and subtract, or fused
        d_sep = a00 * a11 - a10sq
        d_fma = (a00.astype(np.longdouble) * a11.astype(np.longdouble) - a10sq.astype(np.longdouble)).astype(dtype)
        for d in (d_sep, d_fma):
            sure = (a00 > 0) & (d > margin * a10sq)
            assert not np.any(sure & exact_fails)
            assert sure.mean() > 0.1  # the filter does fire on the well-conditioned part


def test_gravity_model_heading_by_angle_addition():
    """step_trig (csrc/cilqr_model.cuh): for the centre-of-gravity model sin / cos of beta = atan(tan(steer) / 2) are
    taken from tan(beta) and the heading (beta + yaw) by angle addition, instead of atan + a third sin/cos as the
    reference does (src/utils.cpp:262-283).  The documented bound: a few 1e-16 absolute on the heading, a few ulp on
    the turn term."""
    rng = np.random.default_rng(11)
    n = 1_000_000
    steer = rng.uniform(-1.2, 1.2, n)
    yaw = rng.uniform(-3.2, 3.2, n)
    # reference sequence
    beta = np.arctan(np.tan(steer) / 2)
    ref_s, ref_c, ref_turn = np.sin(beta + yaw), np.cos(beta + yaw), np.sin(beta)
    # device sequence
    td = np.sin(steer) / np.cos(steer)
    tb = 0.5 * td
    cb = 1.0 / np.sqrt(1.0 + tb * tb)
    sb = tb * cb
    s_head = sb * np.cos(yaw) + cb * np.sin(yaw)
    c_head = cb * np.cos(yaw) - sb * np.sin(yaw)
    assert np.abs(s_head - ref_s).max() < 6e-16 and np.abs(c_head - ref_c).max() < 6e-16
    assert np.abs(sb - ref_turn).max() < 4e-16
    nz = np.abs(ref_turn) > 1e-3
    assert (np.abs(sb - ref_turn)[nz] / np.abs(ref_turn)[nz]).max() < 8 * np.finfo(float).eps
