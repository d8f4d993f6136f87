"""CPU restatement (fp32 torch, no autograd) of the arithmetic the training kernels perform, step by step as the CUDA
code does it (ccdm_b200/csrc/backward.cu, train_kernels.cu, linattn.cu; host sequence of ccdm_b200/train.py).  Used by
tests/test_train_emulated.py to check the DERIVATIONS against torch.autograd without a GPU."""
import torch


def block_bwd_emu(dy, z, gain, gain_mul, ss=None, silu=True):
    """ccdm_block_bwd + ccdm_block_bwd_finish.  dy, z: [B, P, C]; gain [C]; ss [B, 2C] (scale | shift) or None.
    Returns dz, d_ss, dgain, dbias."""
    B, P, C = z.shape
    g = gain * gain_mul
    sc = ss[:, None, :C] if ss is not None else torch.zeros(B, 1, C)
    sh = ss[:, None, C:] if ss is not None else torch.zeros(B, 1, C)
    a = g * (1 + sc)                                            # per (sample, channel)
    inv = torch.rsqrt(z.pow(2).sum(-1, keepdim=True).clamp_min(1e-24))
    zh = z * inv
    du = dy
    if silu:
        u = zh * a + sh
        sig = torch.sigmoid(u)
        du = dy * sig * (1 + u * (1 - sig))
    s0 = (du * zh).sum(1)                                       # sums[0][b][c]
    s1 = du.sum(1)                                              # sums[1]
    dzh = du * a
    dot = (zh * dzh).sum(-1, keepdim=True)
    dz = (dzh - zh * dot) * inv
    s2 = dz.sum(1)                                              # sums[2]
    d_ss = torch.cat([g * s0, s1], dim=1) if ss is not None else None
    dgain = gain_mul * ((1 + sc[:, 0]) * s0).sum(0)
    dbias = s2.sum(0)
    return dz, d_ss, dgain, dbias


def linattn_core_fwd_emu(qkv, scale):
    """ccdm_linattn_prep + ccdm_linattn_context + per-sample block-diagonal GEMM.  qkv [B, n, 384] raw."""
    B, n, _ = qkv.shape
    q, k, v = qkv[..., :128], qkv[..., 128:256], qkv[..., 256:]
    q_sm = q.reshape(B, n, 4, 32).softmax(-1).reshape(B, n, 128) * scale
    p = torch.exp(k - k.max(dim=1, keepdim=True).values)      # shift by the per-sample max over tokens
    S = p.sum(1)                                                # [B, 128]
    ctx = torch.einsum("bnhd,bnhe->bhde", p.reshape(B, n, 4, 32), v.reshape(B, n, 4, 32)) / S.reshape(B, 4, 32, 1)
    out = torch.einsum("bhde,bnhd->bnhe", ctx, q_sm.reshape(B, n, 4, 32)).reshape(B, n, 128)
    return out, (q_sm, p, v, ctx, S)


def linattn_core_bwd_emu(dout, saved, scale):
    """LinAttnCoreFn.backward: dctx (token GEMM), c = rowdot, three per-sample products, finish formulas."""
    q_sm, p, v, ctx, S = saved
    B, n, _ = dout.shape
    r4 = lambda t: t.reshape(B, n, 4, 32)
    dctx = torch.einsum("bnhd,bnhe->bhde", r4(q_sm), r4(dout))                 # ccdm_linattn_dcontext
    G = dctx / S.reshape(B, 4, 32, 1)
    c = (dctx * ctx).sum(-1).reshape(B, 128) / S                                # ccdm_linattn_bwd_rowdot
    dq_sm = torch.einsum("bhde,bnhe->bnhd", ctx, r4(dout)).reshape(B, n, 128)   # per-sample GEMM with ctx
    dp_term = torch.einsum("bhde,bnhe->bnhd", G, r4(v)).reshape(B, n, 128)      # ... with G over v
    dv = torch.einsum("bhde,bnhd->bnhe", G, r4(p)).reshape(B, n, 128)           # ... with G^T over p
    dot = (r4(q_sm) * r4(dq_sm)).sum(-1, keepdim=True) / scale                  # ccdm_linattn_bwd_finish
    dq = (r4(q_sm) * (r4(dq_sm) - dot)).reshape(B, n, 128)
    dk = p * (dp_term - c[:, None, :])
    return torch.cat([dq, dk, dv], -1)


def attention_small_bwd_emu(qkv, dout, heads, dh, scale):
    """ccdm_attention_small_bwd: per (sample, head) recomputation.  qkv [B, n, 3*hid] raw, dout [B, n, hid]."""
    B, n, _ = qkv.shape
    hid = heads * dh
    q, k, v = (qkv[..., i * hid:(i + 1) * hid].reshape(B, n, heads, dh) for i in range(3))
    g = dout.reshape(B, n, heads, dh)
    s = torch.einsum("bihd,bjhd->bhij", q * scale, k)
    P = torch.exp(s - s.max(-1, keepdim=True).values)
    P = P / P.sum(-1, keepdim=True)
    gv = torch.einsum("bihd,bjhd->bhij", g, v)                                  # <dO_i, v_j>
    D = (P * gv).sum(-1, keepdim=True)                                          # <dO_i, O_i>
    dS = P * (gv - D)
    dq = torch.einsum("bhij,bjhd->bihd", dS, k) * scale
    dk = torch.einsum("bhij,bihd->bjhd", dS, q) * scale
    dv = torch.einsum("bhij,bihd->bjhd", P, g)
    return torch.cat([t.reshape(B, n, hid) for t in (dq, dk, dv)], -1)
