"""Experiment drivers: the data side on CPU, one whole experiment on the GPU."""
import numpy as np
import pytest


def test_datasets_follow_the_script_recipe():
    from mcmc_symreg_b200 import experiments as E
    d = E.make_dataset("sim", seed=3)
    assert d["X"].shape == (100, 2) and d["X_test"].shape == (30, 2) and d["X_extra"].shape == (30, 2)          # codes/simulations.py:64-85
    assert np.abs(d["X"]).max() <= 3 and np.abs(d["X_test"]).max() <= 3 and 3 < np.abs(d["X_extra"]).max() <= 6
    X = d["X"]
    np.testing.assert_allclose(d["y"], 1.35 * X[:, 0] * X[:, 1] + 5.5 * np.sin((X[:, 0] - 1) * (X[:, 1] - 1)))   # codes/simulations.py:71
    d2 = E.make_dataset("sim", seed=3)
    assert np.array_equal(d["X"], d2["X"]) and np.array_equal(d["y_extra"], d2["y_extra"])
    assert sorted(E.TARGETS) == ["f1", "f2", "f3", "f4", "f5", "f6", "sim"]
    for name in E.TARGETS:
        dd = E.make_dataset(name, n_train=17, n_test=5, seed=1)
        assert dd["y"].shape == (17,) and np.all(np.isfinite(dd["y"])) and dd["y_test"].shape == (5,)
    assert E.rmse([1.0, 3.0], [0.0, 1.0]) == pytest.approx(np.sqrt(2.5))


@pytest.mark.gpu
def test_one_experiment_end_to_end():
    from mcmc_symreg_b200 import experiments as E
    out, est, data = E.run_experiment("f2", K=3, MM=96, val=60, seed=5)
    assert out["restarts"] == 96 and len(out["model_best"]) == 3 and len(out["beta_best"]) == 4
    # the summaries are what the estimator's own calls give
    assert out["rmse_train_last"] == pytest.approx(E.rmse(est.predict(data["X"]), data["y"]))
    assert out["rmse_test_best"] == pytest.approx(E.rmse(est.predict_best(data["X_test"]), data["y_test"]))
    assert out["complexity_last"] == est.complexity() and out["model_last"] == est.model()
    # the best restart's training RMSE is the one in its trace, and the best of 96 restarts beats their median
    best = est.best_chain()
    # (the trace holds the RMSE from the Gram of the fp32 columns, predict() evaluates in float64: a restart that recovers the
    # target exactly leaves an RMSE of 1e-6 of the scale of y, where the two differ by rounding)
    assert out["rmse_train_best"] == pytest.approx(est.final_rmse_[best], rel=1e-3, abs=1e-5 * float(np.std(data["y"])))
    assert out["rmse_train_best"] <= np.nanmedian(est.final_rmse_)
    assert out["rmse_train_best"] < 0.25 * float(np.std(data["y"]))
    assert 0 < out["accept_rate"] < 0.2 and out["proposals"] > 96 * 60
    assert np.isfinite(out["rmse_extrapolation_best"]) or np.isfinite(out["rmse_extrapolation_mean"])
    line = E.format_row(dict(out, fit_seconds=1.0))
    assert line.startswith("f2")


@pytest.mark.gpu
def test_command_line_suite():
    from mcmc_symreg_b200 import experiments as E
    rows = E.main(["--func", "all", "--MM", "32", "--val", "40", "--seed", "2"])
    assert [r["func"] for r in rows] == ["f1", "f2", "f3", "f4", "f5", "f6"]
    assert all(np.isfinite(r["rmse_train_best"]) for r in rows)
