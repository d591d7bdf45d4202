"""oracle/oracle.cpp (the fp64 restatement every GPU parity test is checked against) held to OUTPUTS OF THE REFERENCE'S OWN CPU CODE.

tests/golden/golden_ref_cpu.json was produced by /root/reference/src/unified_launcher.hpp and everything it includes, compiled
unmodified against an Eigen-API stand-in (oracle/ref_cpu/; Eigen itself is absent from the image) and run through
UnifiedLauncher<CpuBackend>::train / run_full_batch_cpu by tests/golden/make_golden_ref_cpu.py. Both sides are fp64 and differ
only in summation order, so the tolerances are tight: they leave room for rounding amplified along a trajectory, not for an
algorithmic difference (a different line-search branch, ring slot, RNG draw or batch index shows up at 1e-3 .. 1).

  objective (run_full_batch_cpu closures, src/unified_optimization.hpp:101-120)       loss 1e-13, gradient 1e-12
  L-BFGS weak Wolfe (src/minimizer/lbfgs.hpp:38-139, full_batch_minimizer.hpp:126-157) loss 1e-11 (first 10 its) / 1e-7, params 1e-6
  GD (src/minimizer/gd.hpp:42-69)                                                       loss 1e-12, params 1e-11
  SGD random mini-batches (src/minimizer/s_gd.hpp:63-170)                               params 1e-11, CSV loss to its 6 digits
  S-LBFGS (src/minimizer/s_lbfgs.hpp:88-290 + UnifiedSLBFGS_CPU closures)              params after EVERY epoch 1e-8, CSV loss
"""
import json
import os

import numpy as np
import pytest

from golden_cases import ACTS, digest_close, case_problem

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_ref_cpu.json")))["cases"]
IDS = ["-".join(map(str, g["case"]["dims"])) for g in GOLD]


def _close(v, d, rtol):
    err, tol = digest_close(v, d, rtol)
    assert err <= tol, (err, tol)


def _csv_close(vals, csv_vals):
    # the strategies write their history with ostream's default 6 significant digits
    assert len(vals) == len(csv_vals)
    for a, b in zip(vals, csv_vals):
        assert abs(a - b) <= 6e-6 * abs(b), (a, b)


@pytest.mark.parametrize("g", GOLD, ids=IDS)
def test_objective_and_init(oracle, g):
    case = g["case"]
    dims = case["dims"]
    net = oracle.OracleNet(dims, ACTS[tuple(dims)])
    _close(net.init_params_cpu_rule(123), g["bind_params"], 1e-14)  # Network::bindParams (src/network.hpp:45-74)
    w0, X, T = case_problem(case)
    loss, grad = net.loss_grad(w0, X, T)
    assert abs(loss - g["objective"]["loss"]) <= 1e-13 * abs(loss)
    _close(grad, g["objective"]["grad"], 1e-12)


@pytest.mark.parametrize("g", GOLD, ids=IDS)
def test_lbfgs_wolfe_trajectory(oracle, g):
    case = g["case"]
    dims = case["dims"]
    net = oracle.OracleNet(dims, ACTS[tuple(dims)])
    w0, X, T = case_problem(case)
    r = net.lbfgs(w0, X, T, m=case["m"], max_iters=case["lbfgs_iters"], tol=0.0, policy="cpu")
    ref = g["lbfgs"]
    assert r["iters"] == ref["iters"] == case["lbfgs_iters"]
    np.testing.assert_allclose(r["loss"][:10], ref["loss"][:10], rtol=1e-11)  # before rounding has had iterations to grow
    np.testing.assert_allclose(r["loss"], ref["loss"], rtol=1e-7)
    np.testing.assert_allclose(r["gnorm"], ref["gnorm"], rtol=1e-5)
    _close(r["params"], ref["params"], 1e-6)


@pytest.mark.parametrize("g", GOLD, ids=IDS)
def test_gd_trajectory(oracle, g):
    case = g["case"]
    dims = case["dims"]
    net = oracle.OracleNet(dims, ACTS[tuple(dims)])
    w0, X, T = case_problem(case)
    r = net.gd(w0, X, T, lr=case["gd_lr"], momentum=0.0, max_iters=case["gd_iters"], tol=0.0, policy="cpu")
    np.testing.assert_allclose(r["loss"], g["gd"]["loss"], rtol=1e-12)
    _close(r["params"], g["gd"]["params"], 1e-11)


@pytest.mark.parametrize("g", GOLD, ids=IDS)
def test_sgd_random_batches(oracle, g):
    case = g["case"]
    dims = case["dims"]
    net = oracle.OracleNet(dims, ACTS[tuple(dims)])
    w0, X, T = case_problem(case)
    r = net.sgd_cpu_policy(w0, X, T, batch_size=case["sgd_batch"], lr=case["sgd_lr"], max_iters=case["sgd_epochs"])
    _close(r["params"], g["sgd"]["params"], 1e-11)
    _csv_close(r["loss"], g["sgd"]["loss_csv"])
    _csv_close(r["gnorm"], g["sgd"]["gnorm_csv"])


@pytest.mark.parametrize("g", GOLD, ids=IDS)
def test_slbfgs_every_epoch(oracle, g):
    case = g["case"]
    dims = case["dims"]
    net = oracle.OracleNet(dims, ACTS[tuple(dims)])
    w0, X, T = case_problem(case)
    for ref in g["slbfgs"]:
        o = ref["opts"]
        for ep in range(1, o["epochs"] + 1):
            r = net.slbfgs(w0, X, T, batch_size=o["batch"], M=o["M"], L=o["L"], b_H=o["b_H"], step=o["step"], max_iters=ep, tol=0.0)
            _close(r["params"], ref["params_after_epoch"][ep - 1], 1e-8)
        _csv_close(r["loss"], ref["loss_csv"])
        _csv_close(r["gnorm"], ref["gnorm_csv"])


def test_goldens_regenerate_from_the_reference_sources(oracle):
    """where the reference build travels (oracle/_ref/libref_cpu.so): the committed goldens are what it produces today"""
    from oracle import ref_cpu_binding as rc
    if not rc.available():
        pytest.skip("oracle/_ref/libref_cpu.so not built (needs /root/reference)")
    g = GOLD[0]
    case = g["case"]
    net = rc.RefCpuNet(case["dims"])
    w0, X, T = case_problem(case)
    loss, grad = net.loss_grad(w0, X, T)
    assert abs(loss - g["objective"]["loss"]) <= 1e-13 * abs(loss)  # (the summation order of the stand-in's GEMM depends on the thread count)
    _close(grad, g["objective"]["grad"], 1e-12)
    r = net.full_batch("lbfgs", w0, X, T, max_iters=case["lbfgs_iters"], tolerance=0.0, m_param=case["m"])
    np.testing.assert_allclose(r["loss"], g["lbfgs"]["loss"], rtol=1e-7)
