"""Kernel-level parity (run with -m gpu on a B200): each kernel, called through the C ABI (torch custom ops ->
libgic_b200.so), against a plain fp32 PyTorch / numpy expression of the same op."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import captioner as oc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ops():
    from gpt2_image_captioning_b200 import ops, _capi
    return ops, _capi


def _epi_ref(y, epi, res=None):
    if epi == 1:
        return torch.tanh(y)
    if epi == 2:
        return oc.gelu_new(y)
    if epi == 3:
        return torch.relu(y)
    if epi == 4:
        return y + res
    return y


SHAPES = [(64, 2304, 768), (1, 768, 768), (130, 1000, 512), (300, 3072, 768), (77, 50257, 128), (1024, 768, 3072), (8, 256, 64)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("epi", [0, 1, 2, 3, 4])
def test_sgemm_fp32(M, N, K, epi):
    ops, capi = _ops()
    if epi and N > 4000:
        pytest.skip("epilogues covered on the smaller shapes")
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N + K + epi)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.05).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    C = res.clone() if epi == 4 else torch.empty(M, N, device=DEV)
    ops.test_gemm(capi.DTYPE_F32, A, W, b, C, epi)
    ref = _epi_ref(A.double() @ W.double().t() + b.double(), epi, res.double())
    err = (C.double() - ref).abs().max().item()
    assert err <= 2e-5 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


@pytest.mark.parametrize("M,N,K", SHAPES + [(1024, 2304, 768), (10240, 768, 768), (129, 40, 3840)])
@pytest.mark.parametrize("mode", ["bf16", "bf16x2"])
def test_gemm_tcgen05(M, N, K, mode):
    """bf16: against the same product with both operands rounded to bf16 (fp64 accumulate) -> only accumulation
    order differs.  bf16x2 (hi+lo split, 3 MMAs): against the exact fp64 product, ~2^-16 relative per term."""
    ops, capi = _ops()
    g = torch.Generator(device="cpu").manual_seed(M + 3 * N + 5 * K)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.05).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    C = torch.empty(M, N, device=DEV)
    dt = capi.DTYPE_BF16 if mode == "bf16" else capi.DTYPE_BF16X2
    ops.test_gemm(dt, A, W, b, C, 0)
    torch.cuda.synchronize()
    if mode == "bf16":
        ref = A.bfloat16().double() @ W.bfloat16().double().t() + b.double()
        tol = 2e-5
    else:
        ref = A.double() @ W.double().t() + b.double()
        tol = 6e-5
    scale = (A.double().abs() @ W.double().abs().t()).max().item()  # error scales with sum |a||w|
    err = (C.double() - ref).abs().max().item()
    assert err <= tol * scale, f"{mode} M={M} N={N} K={K}: max abs err {err} (scale {scale})"


@pytest.mark.parametrize("M,N,K", [(256, 64, 64), (256, 256, 128), (512, 2304, 768), (1024, 768, 3072), (1000, 3072, 768), (2048, 50272, 64), (10240, 768, 768)])
def test_gemm_cta_pair(monkeypatch, M, N, K):
    """cta_group::2: two CTAs share a 256-row tile (each its own A rows and half of the W tile).  GIC_GEMM_PAIR=2 forces the pair
    kernel wherever the shape allows it; the result must equal the single-CTA kernel's bit for bit (same products, same order)."""
    ops, capi = _ops()
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.05).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    C1 = torch.empty(M, N, device=DEV)
    C2 = torch.empty(M, N, device=DEV)
    monkeypatch.setenv("GIC_GEMM_PAIR", "0")
    ops.test_gemm(capi.DTYPE_BF16, A, W, b, C1, 0)
    monkeypatch.setenv("GIC_GEMM_PAIR", "2")
    ops.test_gemm(capi.DTYPE_BF16, A, W, b, C2, 0)
    torch.cuda.synchronize()
    ref = A.bfloat16().double() @ W.bfloat16().double().t() + b.double()
    scale = (A.double().abs() @ W.double().abs().t()).max().item()
    err = (C2.double() - ref).abs().max().item()
    assert err <= 2e-5 * scale, f"pair M={M} N={N} K={K}: max abs err {err} (scale {scale})"
    assert torch.equal(C1, C2)


@pytest.mark.parametrize("M,d", [(1024, 768), (388, 768), (256, 128), (1300, 256)])
def test_fused_ln_mlp_block_cta_pair(monkeypatch, M, d):
    """the fused MLP sub-block with both GEMMs forced onto CTA pairs (where the tile count is even): same bits as single CTAs"""
    ops, capi = _ops()
    g = torch.Generator(device="cpu").manual_seed(M + d)
    h0 = torch.randn(M, d, generator=g) * 1.5 + 0.3
    gamma = 1.0 + 0.1 * torch.randn(d, generator=g)
    beta = 0.1 * torch.randn(d, generator=g)
    wfc = torch.randn(4 * d, d, generator=g) * 0.03
    bfc = torch.randn(4 * d, generator=g) * 0.1
    wfc2 = torch.randn(d, 4 * d, generator=g) * 0.03
    bfc2 = torch.randn(d, generator=g) * 0.1
    outs = []
    for mode in ("0", "2"):
        monkeypatch.setenv("GIC_GEMM_PAIR", mode)
        h = h0.clone().to(DEV)
        hb = torch.empty(M, d, dtype=torch.bfloat16, device=DEV)
        stats = torch.zeros(d // 32, M, 2, device=DEV)
        ops.test_ln_mlp(h, gamma.to(DEV), beta.to(DEV), wfc.to(DEV), bfc.to(DEV), wfc2.to(DEV), bfc2.to(DEV), hb, stats, 1)
        torch.cuda.synchronize()
        outs.append((h.cpu(), hb.cpu(), stats.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])


@pytest.mark.parametrize("epi", [1, 2, 3, 4])
def test_gemm_tcgen05_epilogues(epi):
    ops, capi = _ops()
    M, N, K = 200, 384, 256
    g = torch.Generator(device="cpu").manual_seed(epi)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.05).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    C = res.clone() if epi == 4 else torch.empty(M, N, device=DEV)
    ops.test_gemm(capi.DTYPE_BF16X2, A, W, b, C, epi)
    ref = _epi_ref(A.double() @ W.double().t() + b.double(), epi, res.double())
    assert (C.double() - ref).abs().max().item() <= 2e-4


@pytest.mark.parametrize("M,d,split_k", [(1024, 768, 1), (1024, 768, 3), (388, 768, 3), (70, 128, 2), (200, 1024, 4), (1300, 256, 1)])
def test_fused_ln_mlp_block(M, d, split_k):
    """ln_2 folded into c_fc (row statistics in, gamma / beta in the packed weight and bias), GELU, c_proj + residual with the
    bf16 copy and the per-32-column row statistics of the new residual stream, fc2 optionally K-split: against the same
    block in fp64 on the bf16-rounded operands (HF GPT2Block ln_2 + GPT2MLP)."""
    ops, capi = _ops()
    g = torch.Generator(device="cpu").manual_seed(M + d + split_k)
    h0 = torch.randn(M, d, generator=g) * 1.5 + 0.3
    gamma = 1.0 + 0.1 * torch.randn(d, generator=g)
    beta = 0.1 * torch.randn(d, generator=g)
    wfc = torch.randn(4 * d, d, generator=g) * 0.03
    bfc = torch.randn(4 * d, generator=g) * 0.1
    wfc2 = torch.randn(d, 4 * d, generator=g) * 0.03
    bfc2 = torch.randn(d, generator=g) * 0.1
    h = h0.clone().to(DEV)
    hb = torch.empty(M, d, dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(d // 32, M, 2, device=DEV)
    ops.test_ln_mlp(h, gamma.to(DEV), beta.to(DEV), wfc.to(DEV), bfc.to(DEV), wfc2.to(DEV), bfc2.to(DEV), hb, stats, split_k)
    torch.cuda.synchronize()
    # reference with the engine's rounding points: xb = bf16(h); W' = bf16(gamma * Wfc); f = bf16(gelu(...)); W2 = bf16(Wfc2)
    xb = h0.bfloat16().double()
    mu = xb.mean(dim=1, keepdim=True)
    var = (xb * xb).mean(dim=1, keepdim=True) - mu * mu
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    w1 = (wfc * gamma).bfloat16().double()
    pre = rstd * (xb @ w1.t() - mu * w1.sum(dim=1)) + (bfc.double() + wfc.double() @ beta.double())
    f = oc.gelu_new(pre.float()).bfloat16().double()
    ref = h0.double() + f @ wfc2.bfloat16().double().t() + bfc2.double()
    err = (h.cpu().double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 4e-3 * scale, f"fused block: max abs err {err} (scale {scale})"  # one bf16 ulp of the GELU output here and there
    # the bf16 copy and its statistics are exact functions of what was written
    assert torch.equal(hb.cpu(), h.cpu().bfloat16())
    hbf = hb.cpu().float().double().view(M, d // 32, 32)
    want = torch.stack((hbf.sum(dim=2), (hbf * hbf).sum(dim=2)), dim=2).permute(1, 0, 2)  # [d/32][M][2]
    got = stats.cpu().double()
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-4)
    # split-K sums the partials in a fixed order: bit-identical across runs
    h2 = h0.clone().to(DEV)
    ops.test_ln_mlp(h2, gamma.to(DEV), beta.to(DEV), wfc.to(DEV), bfc.to(DEV), wfc2.to(DEV), bfc2.to(DEV), hb, stats, split_k)
    torch.cuda.synchronize()
    assert torch.equal(h2, h)


@pytest.mark.parametrize("variant", [-1, 0, 1, 10, 11, 12, 14])
@pytest.mark.parametrize("rows,H,pos,t_max", [(3, 12, 0, 8), (5, 12, 1, 40), (64, 12, 15, 40), (33, 16, 16, 40), (130, 12, 38, 40), (17, 20, 69, 72), (300, 12, 31, 40)])
def test_attn_decode_kernel(variant, rows, H, pos, t_max):
    """One query per (row, head) against `pos` cached keys + its own: softmax(q K^T / 8) V in fp64 on the same bf16 inputs;
    the new K / V must land at cache position `pos` and nothing else in the cache may change."""
    ops, _ = _ops()
    g = torch.Generator(device="cpu").manual_seed(rows * 1000 + pos)
    d = H * 64
    qkv = (torch.randn(rows, 3 * d, generator=g) * 1.5).bfloat16()
    kc = torch.randn(rows, H, t_max, 64, generator=g).bfloat16()
    vc = torch.randn(rows, H, t_max, 64, generator=g).bfloat16()
    kc[:, :, pos:] = float("nan")  # slots that hold no token yet must never be read
    vc[:, :, pos:] = float("nan")
    out = torch.zeros(rows, d, dtype=torch.bfloat16, device=DEV)
    kd, vd = kc.to(DEV), vc.to(DEV)
    ops.test_attn_decode(qkv.to(DEV), kd, vd, out, pos, variant)
    torch.cuda.synchronize()
    q = qkv[:, :d].double().view(rows, H, 1, 64)
    kn = qkv[:, d:2 * d].view(rows, H, 1, 64)
    vn = qkv[:, 2 * d:].view(rows, H, 1, 64)
    K = torch.cat((kc[:, :, :pos], kn), dim=2).double()
    V = torch.cat((vc[:, :, :pos], vn), dim=2).double()
    p = torch.softmax((q @ K.transpose(2, 3)) / 8.0, dim=-1)
    ref = (p @ V).view(rows, d)
    err = (out.cpu().double() - ref).abs().max().item()
    assert err <= 2.0 ** -8 * max(1.0, ref.abs().max().item()), f"max abs err {err}"
    kc2, vc2 = kc.clone(), vc.clone()
    kc2[:, :, pos] = kn[:, :, 0]
    vc2[:, :, pos] = vn[:, :, 0]
    same = lambda a, b: torch.equal(a.view(torch.int16), b.view(torch.int16))  # bit-exact, NaN slots included
    assert same(kd.cpu(), kc2) and same(vd.cpu(), vc2)


@pytest.mark.parametrize("shared", [True, False])
@pytest.mark.parametrize("f16", [False, True])
@pytest.mark.parametrize("images,beams,H,n_prefix,pos,t_max", [(3, 5, 16, 40, 40, 72), (7, 5, 16, 40, 41, 72), (20, 5, 16, 40, 56, 72), (9, 4, 12, 10, 39, 40),
                                                               (5, 8, 12, 17, 34, 48), (33, 2, 20, 0, 16, 40), (40, 5, 16, 40, 69, 70), (11, 3, 12, 16, 48, 64)])
def test_attn_decode_beam_kernel(images, beams, H, n_prefix, pos, t_max, f16, shared):
    """Beam-search decode attention without a cache reorder: hypothesis r attends to the image's prefix (prefill row of the image), to generated
    position g in the cache row its ancestry table names, and to its own new token -- fp64 softmax(q K^T / 8) V on the same 16-bit inputs.  Both
    kernels (prefix shared by the image's beams / one walk per hypothesis), both element types; the new K / V must land in row r at `pos` and the
    rest of the cache must stay untouched (unwritten slots hold NaN and must never be read)."""
    ops, _ = _ops()
    rows, d, ngen = images * beams, H * 64, pos - n_prefix
    et = torch.float16 if f16 else torch.bfloat16
    g = torch.Generator(device="cpu").manual_seed(rows * 1000 + pos + beams)
    qkv = (torch.randn(rows, 3 * d, generator=g) * 1.5).to(et)
    kc = torch.randn(rows, H, t_max, 64, generator=g).to(et)
    vc = torch.randn(rows, H, t_max, 64, generator=g).to(et)
    kc[:, :, pos:] = float("nan")
    vc[:, :, pos:] = float("nan")
    # generated position g of hypothesis r lives in the row of some hypothesis of the same image
    anc = (torch.arange(rows).view(rows, 1) // beams) * beams + torch.randint(0, beams, (rows, max(ngen, 1)), generator=g)
    anc = anc.int().contiguous()
    as16 = lambda x: x.view(torch.bfloat16) if f16 else x  # the C ABI moves 2-byte elements
    out = torch.zeros(rows, d, dtype=torch.bfloat16, device=DEV)
    out_lo = torch.zeros(rows, d, dtype=torch.bfloat16, device=DEV)
    kd, vd = as16(kc).to(DEV), as16(vc).to(DEV)
    ops.test_attn_decode_beam(as16(qkv).to(DEV), kd, vd, out, out_lo, anc.to(DEV), pos, n_prefix, beams, f16, shared)
    torch.cuda.synchronize()
    q = qkv[:, :d].double().view(rows, H, 1, 64)
    kn = qkv[:, d:2 * d].view(rows, H, 1, 64)
    vn = qkv[:, 2 * d:].view(rows, H, 1, 64)
    img_row = (torch.arange(rows) // beams) * beams
    parts_k, parts_v = [kc[img_row][:, :, :n_prefix]], [vc[img_row][:, :, :n_prefix]]
    for gi in range(ngen):
        src = anc[:, gi].long()
        parts_k.append(kc[src][:, :, n_prefix + gi:n_prefix + gi + 1])
        parts_v.append(vc[src][:, :, n_prefix + gi:n_prefix + gi + 1])
    K = torch.cat(parts_k + [kn], dim=2).double()
    V = torch.cat(parts_v + [vn], dim=2).double()
    p = torch.softmax((q @ K.transpose(2, 3)) / 8.0, dim=-1)
    ref = (p @ V).view(rows, d)
    got = out.cpu().double() + (out_lo.cpu().double() if f16 else 0.0)
    err = (got - ref).abs().max().item()
    tol = (2.0 ** -11 if f16 else 2.0 ** -8) * max(1.0, ref.abs().max().item())
    assert err <= tol, f"max abs err {err} (tol {tol})"
    kc2, vc2 = kc.clone(), vc.clone()
    kc2[:, :, pos] = kn[:, :, 0]
    vc2[:, :, pos] = vn[:, :, 0]
    same = lambda a, b: torch.equal(a.view(torch.int16), b.view(torch.int16))  # bit-exact, NaN slots included
    assert same(kd.cpu(), as16(kc2)) and same(vd.cpu(), as16(vc2))


@pytest.mark.parametrize("rows,H,S,t_max", [(3, 12, 1, 8), (37, 12, 10, 40), (9, 16, 16, 20), (11, 16, 17, 72), (6, 12, 40, 70), (5, 20, 64, 64), (2, 12, 70, 80)])
def test_attn_prefill_kernel(rows, H, S, t_max):
    """causal attention over the S prefix tokens of every (row, head) in fp64 on the same bf16 inputs; K / V of the S tokens
    must land in the cache and the rest of the cache must stay untouched (S = 70 takes the generic kernel)."""
    ops, _ = _ops()
    g = torch.Generator(device="cpu").manual_seed(rows * 100 + S)
    d = H * 64
    qkv = (torch.randn(rows * S, 3 * d, generator=g) * 1.5).bfloat16()
    kc = torch.full((rows, H, t_max, 64), float("nan")).bfloat16()
    vc = torch.full((rows, H, t_max, 64), float("nan")).bfloat16()
    out = torch.zeros(rows * S, d, dtype=torch.bfloat16, device=DEV)
    kd, vd = kc.to(DEV), vc.to(DEV)
    ops.test_attn_prefill(qkv.to(DEV), kd, vd, out, S)
    torch.cuda.synchronize()
    q, k, v = (qkv[:, i * d:(i + 1) * d].view(rows, S, H, 64).permute(0, 2, 1, 3) for i in range(3))
    sc = (q.double() @ k.double().transpose(2, 3)) / 8.0
    sc = sc.masked_fill(torch.triu(torch.ones(S, S, dtype=torch.bool), diagonal=1), float("-inf"))
    ref = (torch.softmax(sc, dim=-1) @ v.double()).permute(0, 2, 1, 3).reshape(rows * S, d)
    err = (out.cpu().double() - ref).abs().max().item()
    assert err <= 2.0 ** -8 * max(1.0, ref.abs().max().item()), f"max abs err {err}"
    kc[:, :, :S] = k
    vc[:, :, :S] = v
    same = lambda a, b: torch.equal(a.contiguous().view(torch.int16), b.contiguous().view(torch.int16))
    assert same(kd.cpu(), kc) and same(vd.cpu(), vc)


@pytest.mark.parametrize("temperature,top_p", [(1.0, 0.9), (0.7, 0.5), (1.5, 1.0), (1.0, 0.05)])
def test_sample_top_p_kernel_distribution(temperature, top_p):
    """40 000 draws from one logit row: every token lies in the reference's nucleus (its own torch expression, src/models.py:413-432)
    and the empirical distribution matches the renormalised softmax (total variation); a second row with a different nucleus
    is sampled in the same launch."""
    ops, _ = _ops()
    V, n = 3000, 40000
    g = torch.Generator(device="cpu").manual_seed(int(temperature * 100 + top_p * 1000))
    rows = torch.stack((torch.randn(V, generator=g) * 2.0, torch.randn(V, generator=g) * 4.0))
    which = torch.arange(n) % 2
    logits = rows[which].to(DEV).contiguous()
    tok = torch.empty(n, dtype=torch.int32, device=DEV)
    ops.test_sample_top_p(logits, temperature, top_p, 1234, 3, tok)
    torch.cuda.synchronize()
    tok2 = torch.empty_like(tok)
    ops.test_sample_top_p(logits, temperature, top_p, 1234, 3, tok2)  # same (seed, row, step) -> same draws
    ops.test_sample_top_p(logits, temperature, top_p, 1234, 4, tok)  # (keep tok2 as the reference draw; another step differs)
    torch.cuda.synchronize()
    assert not torch.equal(tok, tok2)
    ops.test_sample_top_p(logits, temperature, top_p, 1234, 3, tok)
    torch.cuda.synchronize()
    assert torch.equal(tok, tok2)
    tok = tok.cpu().long()
    for r in range(2):
        z = rows[r:r + 1] / temperature
        if top_p < 1.0:  # the reference's nucleus mask
            sl, si = torch.sort(z, descending=True)
            cp = torch.cumsum(torch.softmax(sl, dim=-1), dim=-1)
            rem = cp > top_p
            rem[:, 1:] = rem[:, :-1].clone()
            rem[:, 0] = False
            z = z.masked_fill(rem.scatter(1, si, rem), float("-inf"))
        want = torch.softmax(z, dim=-1)[0].double()
        draws = tok[which == r]
        assert bool((want[draws] > 0).all()), "a token outside the nucleus was drawn"
        emp = torch.bincount(draws, minlength=V).double() / draws.numel()
        tv = 0.5 * (emp - want).abs().sum().item()
        kept = int((want > 0).sum())
        assert tv < 0.02 + 0.6 * (kept / draws.numel()) ** 0.5, (r, tv, kept)


@pytest.mark.parametrize("rows,d", [(1, 768), (1000, 768), (33, 1024), (7, 1280), (5, 128)])
def test_layernorm(rows, d):
    ops, _ = _ops()
    g = torch.Generator(device="cpu").manual_seed(rows + d)
    x = (torch.randn(rows, d, generator=g) * 3 + 1).to(DEV)
    w = torch.randn(d, generator=g).to(DEV)
    b = torch.randn(d, generator=g).to(DEV)
    y = torch.empty_like(x)
    ops.test_layernorm(x, w, b, y)
    ref = torch.nn.functional.layer_norm(x, (d,), w, b, 1e-5)
    assert (y - ref).abs().max().item() <= 2e-5


@pytest.mark.parametrize("B,N,k", [(37, 5000, 14), (1, 100, 5), (5, 70000, 16), (3, 7, 12), (64, 40000, 5), (2, 33000, 30)])
def test_topk_ip_exact_indices(B, N, k):
    """Indices bit-exact against exact IndexFlatIP semantics (descending, ties -> lowest index), incl. k > N padding,
    duplicate database rows (exact ties) and more than one 32768-row chunk."""
    from gpt2_image_captioning_b200.database import _GpuFlatIndex
    rng = np.random.default_rng(B * 31 + N + k)
    D = 64
    db = rng.standard_normal((N, D)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    if N > 50:
        db[N // 2] = db[3]  # exact duplicate -> tie
        db[N - 1] = db[3]
    q = rng.standard_normal((B, D)).astype(np.float32)
    q[0] = db[3]
    idx = _GpuFlatIndex(torch.from_numpy(db).to(DEV))
    s, i = idx.search(q, k)
    # integer-valued inputs would make scores exact; with real data compare to the fp64 ordering and allow a swap only
    # where the exact scores are closer than fp32 rounding
    exact = q.astype(np.float64) @ db.astype(np.float64).T
    ws, wi = oc.flat_ip_search(db, q, k)
    for b in range(B):
        for j in range(k):
            if i[b, j] != wi[b, j]:
                assert i[b, j] >= 0 and wi[b, j] >= 0
                assert abs(exact[b, i[b, j]] - exact[b, wi[b, j]]) < 1e-6, (b, j, i[b], wi[b])
    valid = wi >= 0
    np.testing.assert_allclose(s[valid], ws[valid], atol=2e-6)
    assert np.all(i[~valid] == -1) and np.all(np.isneginf(s[~valid]))
    if N > 50:  # the three identical rows must come out in index order
        pos = [int(np.where(i[0] == r)[0][0]) for r in (3, N // 2, N - 1) if r in i[0]]
        assert pos == sorted(pos)


@pytest.mark.parametrize("B,N,D,k", [(7, 300, 64, 5), (33, 70000, 128, 15), (1024, 40000, 512, 16), (5, 20, 64, 16)])
def test_topk_ip_tensor_core_path_equals_exact_path(B, N, D, k):
    """bf16x2 tcgen05 candidate scan + exact re-scoring + certificate (gic_topk_ip_tc): scores and indices bit-identical to the
    fp32 scan, with exact ties (duplicate rows), k > N padding, and rows whose certificate fails (forced by lying about the
    database norm: eps becomes huge, so every query takes the exact fix-up scan)."""
    from gpt2_image_captioning_b200.database import _GpuFlatIndex
    rng = np.random.default_rng(B + N + D + k)
    db = rng.standard_normal((N, D)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    if N > 50:
        db[N // 2] = db[3]
        db[N - 1] = db[3]
    q = rng.standard_normal((B, D)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[0] = db[3]
    dbt = torch.from_numpy(db).to(DEV)
    exact = _GpuFlatIndex(dbt, tensor_cores=False)
    tc = _GpuFlatIndex(dbt, tensor_cores=True)
    assert tc.hi is not None and exact.hi is None
    se, ie = exact.search(q, k)
    st, it = tc.search(q, k)
    assert np.array_equal(ie, it) and np.array_equal(se, st)
    tc.norm_max = 1e6  # no certificate can hold: every row goes through the exact fix-up kernel
    sf, i_f = tc.search(q, k)
    assert np.array_equal(ie, i_f) and np.array_equal(se, sf)


def test_topk_ip_integer_scores_bit_exact():
    """Integer-valued vectors make every inner product exact in fp32, so scores AND indices must be bit-identical."""
    from gpt2_image_captioning_b200.database import _GpuFlatIndex
    rng = np.random.default_rng(5)
    db = rng.integers(-3, 4, (9000, 32)).astype(np.float32)
    q = rng.integers(-3, 4, (20, 32)).astype(np.float32)
    s, i = _GpuFlatIndex(torch.from_numpy(db).to(DEV)).search(q, 15)
    ws, wi = oc.flat_ip_search(db, q, 15)
    assert np.array_equal(i, wi) and np.array_equal(s, ws)


def test_retrieval_against_reference_fixture():
    """GpuFlatStore vs what the reference's own faiss_store functions + RetrievalAggregator produced (rat_retrieval.npz)."""
    from gpt2_image_captioning_b200 import GpuFlatStore
    g = gu.load("rat_retrieval")
    rng = np.random.default_rng(int(g["db_seed"]))
    n_img, D = int(g["n_img"]), 512
    img = rng.standard_normal((n_img, D)).astype(np.float32)
    img /= np.linalg.norm(img, axis=1, keepdims=True)
    counts = rng.integers(0, 7, n_img)
    owner = np.repeat(np.arange(n_img), counts)
    cap = rng.standard_normal((len(owner), D)).astype(np.float32)
    names = [f"img_{i:06d}.jpg" for i in range(n_img)]
    store = GpuFlatStore(img, cap, names, [{"filename": names[o], "caption_id": j} for j, o in enumerate(owner)], device=DEV)
    q = torch.from_numpy(g["q"])
    for (k, i) in [(10, 4), (20, 6), (5, 1)]:
        # query 6 is the exact midpoint of images 3 and 4: a mathematical tie that fp32 rounding breaks either way, so its
        # two leading hits (and their caption rows) may legitimately swap; every other query must match bit for bit
        TIE_Q = 6
        rows = store.retrieve_rows(q.to(DEV), top_i=i, top_k=k).cpu().numpy()
        starts = np.concatenate([[0], np.cumsum(counts)])
        _, want_rows = oc.retrieve_and_aggregate(img, cap, lambda im: list(range(starts[im], starts[im + 1])), g["q"], top_i=i, top_k=k)
        others = np.arange(rows.shape[0]) != TIE_Q
        assert np.array_equal(rows[others], want_rows[others]), np.nonzero((rows != want_rows).any(axis=1))[0]
        if i > 1:
            assert sorted(rows[TIE_Q].tolist()) == sorted(want_rows[TIE_Q].tolist())
        ret = store.retrieve_caption_embeddings(q.to(DEV), top_i=i, top_k=k).cpu().numpy()
        assert np.array_equal(ret[others], g[f"ret_k{k}_i{i}"][others]), (k, i)
        aug = store.retrieve_and_aggregate(q, top_i=i, top_k=k, aggregation="mean")
        assert aug.device.type == "cpu"
        np.testing.assert_allclose(aug.numpy()[others], g[f"aug_k{k}_i{i}"][others], atol=1e-6)
        if i > 1:  # the mean does not depend on the order of the tied hits
            np.testing.assert_allclose(aug.numpy()[TIE_Q], g[f"aug_k{k}_i{i}"][TIE_Q], atol=1e-6)
    # other pooling modes against the PyTorch expression of RetrievalAggregator (src/models.py:593-606)
    ret = torch.from_numpy(g["ret_k10_i4"])
    want_max = q + ret.max(dim=1)[0]
    want_sn = q + torch.nn.functional.normalize(torch.nn.functional.normalize(ret, p=2, dim=2).sum(dim=1), p=2, dim=1)
    np.testing.assert_allclose(store.retrieve_and_aggregate(q, 4, 10, "max").numpy(), want_max.numpy(), atol=1e-6)
    np.testing.assert_allclose(store.retrieve_and_aggregate(q, 4, 10, "sum_norm").numpy(), want_sn.numpy(), atol=2e-6)
    proj = torch.nn.Linear(512, 1)
    torch.nn.init.normal_(proj.weight, std=0.5)
    want_at = q + (ret * torch.softmax(proj(ret), dim=1)).sum(dim=1)
    got_at = store.retrieve_and_aggregate(q, 4, 10, "attention", attention_weight=proj.weight, attention_bias=proj.bias)
    # (scores w . r are O(10) here -- unnormalised rows, std-0.5 weights -- so fp32 summation order moves the softmax weights by ~1e-6)
    np.testing.assert_allclose(got_at.numpy(), want_at.detach().numpy(), atol=2e-5)
    # duck-typed faiss surface
    s, ix = store.image_index.search(g["q"][:3], 14)
    assert s.shape == (3, 14) and ix.dtype == np.int64
    assert np.array_equal(store.caption_index.reconstruct(7), cap[7])


def test_store_from_faiss_directory_and_embedding_files(tmp_path):
    """GpuFlatStore.from_directory / .from_embedding_files / .save: the store read back from the
    reference's on-disk layout answers exactly like the one built from the matrices."""
    from gpt2_image_captioning_b200 import GpuFlatStore, faiss_files
    rng = np.random.default_rng(11)
    n_img, D = 300, 64
    img = rng.standard_normal((n_img, D)).astype(np.float32)
    img /= np.linalg.norm(img, axis=1, keepdims=True)
    owner = np.repeat(np.arange(n_img), rng.integers(0, 6, n_img))
    cap = rng.standard_normal((len(owner), D)).astype(np.float32)
    names = [f"img_{i:06d}.jpg" for i in range(n_img)]
    meta = [{"filename": names[o], "caption_id": int(j)} for j, o in enumerate(owner)]
    a = GpuFlatStore(img, cap, names, meta, device=DEV)
    assert GpuFlatStore.from_directory(str(tmp_path / "nothing_here"), device=DEV) is None
    a.save(str(tmp_path / "db"))
    b = GpuFlatStore.from_directory(str(tmp_path / "db"), device=DEV)
    entries = [{"filenames": names[i], "embeddings": [{"embedding": torch.from_numpy(cap[j]), "caption_id": int(j)}
                                                      for j in np.nonzero(owner == i)[0]]} for i in range(n_img)]
    torch.save({"filenames": names, "embeddings": torch.from_numpy(img)}, tmp_path / "img.pt")
    torch.save(entries, tmp_path / "cap.pt")
    c = GpuFlatStore.from_embedding_files(str(tmp_path / "img.pt"), str(tmp_path / "cap.pt"), device=DEV)
    q = torch.from_numpy(rng.standard_normal((17, D)).astype(np.float32)).to(DEV)
    want_rows = a.retrieve_rows(q, top_i=4, top_k=10)
    want_aug = a.retrieve_and_aggregate(q, 4, 10, "mean")
    for other in (b, c):
        assert other.image_metadata == names and other.caption_metadata == meta
        assert torch.equal(other.image_index.matrix, a.image_index.matrix) and torch.equal(other.caption_index.matrix, a.caption_index.matrix)
        assert torch.equal(other.retrieve_rows(q, top_i=4, top_k=10), want_rows)
        assert torch.equal(other.retrieve_and_aggregate(q, 4, 10, "mean"), want_aug)
    assert faiss_files.locate_vectors(str(tmp_path / "db" / "image_index.faiss"))[:2] == (D, n_img)
