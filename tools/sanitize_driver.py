"""Small hot-path invocations for compute-sanitizer (tools/gpu_sanitize.sh): every kernel family once, at sizes that stay
fast under the tool.  `python tools/sanitize_driver.py [which ...]`"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from conftest import load_case  # noqa: E402
from test_gpu_parity import run_case_on_gpu  # noqa: E402
from test_gpu_properties import sub_case  # noqa: E402


def case(name, t=20):
    c = load_case(name)
    return sub_case(c, list(range(c["u"].shape[0])), None, t)


def main(which):
    if "dr_latency" in which:  # team prologue forward; matrix-form reverse sweep (mbarrier rings, vh_bwd_mx.cuh)
        run_case_on_gpu(case("dr_constant_icml_midpoint_f32_iw8"))
    if "dr_ws" in which:  # the producer / consumer reverse sweep (named-barrier ring) the matrix form replaced for this model
        os.environ["VIHDS_BWD_MX"] = "0"
        run_case_on_gpu(case("dr_constant_icml_midpoint_f32_iw8"))
        del os.environ["VIHDS_BWD_MX"]
    if "dr_throughput" in which:
        os.environ["VIHDS_FWD_TEAM"], os.environ["VIHDS_BWD_WS"] = "0", "0"
        run_case_on_gpu(case("dr_constant_icml_midpoint_f32_iw8"))
        del os.environ["VIHDS_FWD_TEAM"], os.environ["VIHDS_BWD_WS"]
    if "relay_precisions" in which:  # + NeuralPrecisions weights in shared memory, weight-gradient warp
        run_case_on_gpu(case("relay_constant_precisions_midpoint_f32_iw8"))
    if "hidden_precisions" in which:
        run_case_on_gpu(case("dr_constant_precisions_hidden5_midpoint_f32_iw8"))
    if "blackbox_mma" in which:  # warp-level tensor-core kernels, panel ring between the adjoint and weight-gradient warps
        os.environ["VIHDS_BB_IMPL"] = "mma"
        run_case_on_gpu(case("dr_blackbox_icml_midpoint_f32_iw8", 12))
        c = load_case("dr_blackbox_icml_midpoint_f32_iw8")
        run_case_on_gpu(sub_case(c, [0, 1, 2, 3, 4], 7, 9))  # ragged: 35 trajectories
    if "blackbox_scalar" in which:
        os.environ["VIHDS_BB_IMPL"] = "scalar"
        run_case_on_gpu(case("dr_blackbox_icml_midpoint_f32_iw8", 12))
        os.environ["VIHDS_BB_IMPL"] = "mma"
    if "exchange" in which:  # one-rank gradient exchange + Adam, plain device Adam with the NaN guard
        import test_gpu_properties as T

        T.test_fused_allreduce_adam_single_rank_equals_adam()
        T.test_adam_dev_nan_cost_leaves_parameters_and_moments_untouched()
    if "step" in which:  # the whole graphed training step, eager (encoder kernels, conditioner, iwae, adam)
        from test_gpu_package import build
        from vihds_b200.datasets import batch_of
        from vihds_b200.training import GraphedStep

        settings, par, model, training = build("dr_constant_icml")
        ds = training.dataset_pair.train.dataset
        batch = batch_of(ds, np.asarray(training.dataset_pair.train.indices)[:6], settings.device, settings.dtype)
        model.want_predict = False
        gs = GraphedStep(training, 6, 8, batch.times.numel(), use_graphs=False)
        gs.load_batch(batch)
        gs.load_u(torch.randn(6, 8, par.n_theta, device="cuda"))
        gs.draw_conditioner()
        print("cost", float(gs.step().item()))
    torch.cuda.synchronize()
    print("SANITIZE_DRIVER_DONE", " ".join(which))


if __name__ == "__main__":
    main(sys.argv[1:] or ["dr_latency", "dr_throughput", "relay_precisions", "hidden_precisions", "blackbox_mma", "blackbox_scalar",
                          "exchange", "step"])
