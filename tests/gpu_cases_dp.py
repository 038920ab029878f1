"""Single-GPU emulation of the data-parallel layout: two molecule shards run one after the other on one device, the
3-D embeddings are concatenated (what all_gather_rows produces), each shard's loss rows use (row_offset, total_rows),
and the summed shard losses / gradients must equal the full-batch result when BatchNorm is in eval mode (statistics
independent of the sharding).  The NCCL version of the same check is tests/gpu_dist_check.py (needs 2 GPUs).

Tolerance note: with seeded "trained-scale" running statistics the eval-mode embeddings of different molecules are
almost collinear (loss = ln(B-1) to 5 digits), so d loss / d z suffers catastrophic cancellation in fp32 on ANY
implementation: the CPU oracle in fp32 differs from itself in fp64 by 4e-3 of the (tiny, 1e-5) gradient scale,
the CUDA path by 3e-3 (measured, gpurun_out/dbg_dp.txt).  Hence 2e-2 here; train-mode gradients agree to 3e-5."""
import importlib

import torch

from oracle import oracle as O

i3d = importlib.import_module("3dinfomax_b200")
cfg = importlib.import_module("3dinfomax_b200.configs")
syn = i3d.synthetic
DEV = "cuda"


def _fresh(st2, st3):
    pna = i3d.PNA(avg_d=1, device=DEV, **cfg.PRETRAIN_QM9_MODEL_PARAMETERS).to(DEV)
    n3 = i3d.Net3D(node_dim=0, edge_dim=1, avg_d=1, **cfg.PRETRAIN_QM9_MODEL3D_PARAMETERS).to(DEV)
    pna.load_state_dict(st2), n3.load_state_dict(st3)
    pna.eval(), n3.eval()
    return pna, n3


def case_sharded_equals_full():
    out = []
    C, per, world = 3, 24, 2
    b = syn.make_batch(77, per * world, conformers=C)
    c2, c3 = O.pna_cfg(**cfg.PRETRAIN_QM9_MODEL_PARAMETERS), O.net3d_cfg(**cfg.PRETRAIN_QM9_MODEL3D_PARAMETERS)
    st2, st3 = O.init_pna_state(c2, 5, True), O.init_net3d_state(c3, 6, True)
    L = i3d.lib.load()
    for backend, tag in ((0, "tensor_core"), (1, "simt")):
        L.i3d_gemm_backend(backend)
        loss_fn = i3d.NTXentMultiplePositives(tau=0.1)
        pna, n3 = _fresh(st2, st3)
        g2, g3 = i3d.batch_from_numpy(b, DEV)
        full = loss_fn(pna(g2), n3(g3))
        full.backward()
        gfull = torch.cat([p.grad.reshape(-1) for p in list(pna.parameters()) + list(n3.parameters())]).clone()
        pna, n3 = _fresh(st2, st3)
        z2s, z3s = [], []
        for r in range(world):
            h2, h3 = i3d.batch_from_numpy(syn.slice_batch(b, r * per, (r + 1) * per), DEV)
            z2s.append(pna(h2))
            z3s.append(n3(h3))
        z3_all = torch.cat(z3s)
        total = sum(loss_fn(z2s[r], z3_all, row_offset=r * per, total_rows=per * world) for r in range(world))
        total.backward()
        gsh = torch.cat([p.grad.reshape(-1) for p in list(pna.parameters()) + list(n3.parameters())])
        scale = gfull.abs().max().item()
        out += [("dp_emulation/%s/loss" % tag, abs(total.item() - full.item()), 1e-5),
                ("dp_emulation/%s/grad_err_over_scale" % tag, (gsh - gfull).abs().max().item() / scale, 2e-2)]
        # oracle on the CPU, full batch, eval mode
        if backend == 1:
            og2, xa, ea, og3, d3 = O.graphs_from_batch(b)
            o2, o3 = O.as_leaf_params(st2), O.as_leaf_params(st3)
            ol = O.ntxent_multiple_positives(O.pna_forward(o2, c2, og2, xa, ea, False),
                                             O.net3d_forward(o3, c3, og3, d3, False), tau=0.1)
            ol.backward()
            go = torch.cat([o2[k].grad.reshape(-1) for k in O.param_keys(o2)] +
                           [o3[k].grad.reshape(-1) for k in O.param_keys(o3)])
            out += [("dp_emulation/oracle/loss", abs(full.item() - ol.item()), 1e-5),
                    ("dp_emulation/oracle/grad_err_over_scale(simt)", (gfull.cpu() - go).abs().max().item() / scale, 2e-2)]
    L.i3d_gemm_backend(0)
    return out
