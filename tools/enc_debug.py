import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
from conftest import load_case
from test_gpu_package import build, batch_from_case, _rel
for case_name, spec, dims in [("dr_constant_icml_midpoint_f32_iw8", "dr_constant_icml", None), ("dr_constant_one_midpoint_f32_iw5", "dr_constant_one", (4, 100, 2, 1))]:
    case = load_case(case_name)
    settings, par, model, training = build(spec, dims)
    enc = model.encoder
    batch = batch_from_case(case)
    mu1, pr1 = enc.q_table(batch); mu0, pr0 = enc.q_table_reference(batch)
    print(spec, "fwd", _rel(mu1.detach().cpu().numpy(), mu0.detach().cpu().numpy()), _rel(pr1.detach().cpu().numpy(), pr0.detach().cpu().numpy()))
    g = torch.Generator(device="cuda").manual_seed(0)
    g_mu, g_pr = torch.randn(mu0.shape, device="cuda", generator=g), torch.randn(mu0.shape, device="cuda", generator=g)
    names = ["conv_w", "conv_b", "lin_w", "lin_b", "local_w", "local_b", "gcond_w", "global_free"]
    params = [(n, p) for n, p in zip(names, enc.fused_parameters()) if p.numel()]
    got = torch.autograd.grad([mu1, pr1], [p for _, p in params], [g_mu, g_pr], allow_unused=True)
    ref = torch.autograd.grad([mu0, pr0], [p for _, p in params], [g_mu, g_pr], allow_unused=True)
    for (n, p), a, b in zip(params, got, ref):
        print("  %-12s %-16s rel %.3e  max|ref| %.3e" % (n, tuple(p.shape), _rel(a.cpu().numpy(), b.cpu().numpy()), float(b.abs().max())))
