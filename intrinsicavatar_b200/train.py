"""Training-mode seam of the fused query: ``SNARFDeformer.deform(pts, model, eval_mode=False)``
(models/deformers/snarf_deformer.py:170-261) as ONE differentiable call.

The reference builds the graph out of ``ForwardDeformer.forward`` (search under ``no_grad`` + the implicit-differentiation
correction, models/deformers/fast_snarf/deformer_torch.py:34-76), ``VolumeSDF`` (tiny-cuda-nn hash grid + MLP,
models/rf/geometry.py:147-172) and ``torch.min`` over the candidate roots, and lets autograd walk it.  Here the forward is
``ia_op_query_train`` and the backward ``ia_op_query_backward`` (csrc/ia_train.cuh); this module only routes their
buffers into autograd.  There is no fallback: without the CUDA library the calls raise.

    sdf, feature, x_c, valid = fused_query(engine, xd, geo_hash, geo_w1, geo_b1, geo_w2, geo_b2, tfs)

``geo_hash`` ... ``geo_b2`` are the tensors the engine's fields were set from (``RenderEngine.set_fields``: the hash table
and the effective -- weight-norm folded -- MLP weights) and ``tfs`` [24,4,4] the bone transforms of ``set_pose``: they
are graph leaves here, the VALUES are the engine's, so call ``set_fields`` / ``set_pose`` after every optimiser step.
Gradients reach them from ``sdf`` and ``feature``; ``x_c`` and ``valid`` are returned without a graph (the reference's
canonical points carry the correction's gradient too -- route it through ``engine.op_deform_backward`` if a loss reads them).
"""
from __future__ import annotations

import torch


class _FusedQuery(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, xd, geo_hash, geo_w1, geo_b1, geo_w2, geo_b2, tfs):
        fwd = engine.op_query_train(xd.detach())
        ctx.engine = engine
        ctx.fwd = {k: fwd[k] for k in ("x_c", "valid", "J_inv")}
        ctx.meta = (geo_hash.shape, geo_hash.dtype, tfs.shape)
        ctx.mark_non_differentiable(fwd["x_c"], fwd["valid"])
        return fwd["sdf"], fwd["feature"], fwd["x_c"], fwd["valid"]

    @staticmethod
    def backward(ctx, g_sdf, g_feature, _g_xc, _g_valid):
        n = ctx.fwd["x_c"].shape[0]
        dev = ctx.fwd["x_c"].device
        # the 13 network outputs at the arg-min root: `feature` IS the output vector (channel 0 = sdf, geometry.py:160-166)
        d_out = torch.zeros(n, 13, device=dev) if g_feature is None else g_feature.to(dev, torch.float32).clone()
        if g_sdf is not None:
            d_out[:, 0] += g_sdf.to(dev, torch.float32)
        g = ctx.engine.op_query_backward(ctx.fwd, d_out)
        hash_shape, hash_dtype, tfs_shape = ctx.meta
        g_tfs = torch.zeros(tfs_shape, device=dev)
        g_tfs[..., :3, :] = g["tfs"].reshape(g_tfs[..., :3, :].shape)
        return (None, None, g["hash"].reshape(hash_shape).to(hash_dtype), g["w1"], g["b1"], g["w2"], g["b2"], g_tfs)


def fused_query(engine, xd, geo_hash, geo_w1, geo_b1, geo_w2, geo_b2, tfs):
    """Differentiable posed point -> (sdf [n], feature [n,13], x_c [n,3], valid [n]); see the module docstring."""
    return _FusedQuery.apply(engine, xd, geo_hash, geo_w1, geo_b1, geo_w2, geo_b2, tfs)
