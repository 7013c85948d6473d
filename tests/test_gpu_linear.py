"""GPU parity of rd_linear through the C-ABI: streaming GEMV (M<=4), tcgen05 tiles (all NT widths, split-K, SwiGLU,
LoRA, residual, bias/act) and the SIMT cross-check kernel, against a plain PyTorch fp32 reference of the same op with
the same rounding points.  Tolerance: one storage-dtype ulp of the result magnitude (different fp32 summation order is
the only difference allowed)."""
import ctypes as C

import pytest
import torch

from radialog_b200 import _lib

pytestmark = pytest.mark.gpu


def ref_linear(x, w, dtype, bias=None, act=0, residual=None, res_mode=1, lora_t=None, lora_b=None, lora_scale=0.0, N=None):
    xf, wf = x.float(), w.float()
    if act == _lib.ACT_SWIGLU:
        g = (xf @ wf[:N].t()).to(dtype)
        u = (xf @ wf[N:].t()).to(dtype)
        return torch.nn.functional.silu(g.float()).to(dtype) * u
    v = xf @ wf.t()
    if bias is not None:
        v = v + bias
    if residual is not None and res_mode == 2:
        v = v + residual.float()
    if act == _lib.ACT_RELU:
        v = torch.relu(v)
    elif act == _lib.ACT_GELU:
        v = torch.nn.functional.gelu(v)
    y = v.to(dtype)
    if residual is not None and res_mode == 2:
        return y
    if lora_t is not None:
        s = (lora_t.float() @ lora_b.float().t()).to(dtype)
        y = y + s * lora_scale
    if residual is not None:
        y = residual + y
    return y


def run_linear(lib, x, w, M, N, K, dtype, algo, bias=None, act=0, residual=None, res_mode=1, lora_t=None, lora_b=None,
               lora_scale=0.0, ws=None):
    out = torch.full((M, N), float("nan"), device=x.device, dtype=dtype)
    e = _lib.Epilogue()
    e.bias_dev = _lib.ptr(bias)
    e.residual_dev = _lib.ptr(residual)
    e.ld_res = N
    e.res_mode = res_mode
    e.act = act
    e.lora_t_dev = _lib.ptr(lora_t)
    e.lora_b_dev = _lib.ptr(lora_b)
    e.lora_r = 0 if lora_t is None else lora_t.shape[1]
    e.lora_scale = lora_scale
    wsp, wsn = (_lib.ptr(ws), ws.numel()) if ws is not None else (None, 0)
    st = lib.rd_linear(_lib.ptr(x), K, _lib.ptr(w), K, _lib.ptr(out), N, M, N, K, C.byref(e), _lib.dtype_code(dtype), algo, wsp, wsn,
                       _lib.current_stream())
    _lib.check(st, "rd_linear")
    torch.cuda.synchronize()
    return out


def assert_close_ulp(out, ref, dtype, what, mag=None):
    """|out-ref| <= 2.5 ulp(magnitude) + 1e-5*max|ref| (the absolute floor covers results that cancel to ~0, where the
    fp32 summation-order difference is larger than an ulp of the tiny result), and bit-equal on the vast majority.
    `mag`: magnitude of the intermediate that carries the rounding (e.g. |residual|+|y| for a two-rounding epilogue)."""
    assert torch.isfinite(out.float()).all(), f"{what}: non-finite / unwritten output"
    eps = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    r = ref.float()
    tol = 2.5 * eps * (r.abs() if mag is None else mag.float().abs()) + 1e-5 * r.abs().max()
    err = (out.float() - r).abs()
    worst = (err - tol).max().item()
    assert worst <= 0, f"{what}: error exceeds tolerance by {worst:.3e} (max abs err {err.max().item():.3e})"
    frac_exact = (out == ref).float().mean().item()
    assert frac_exact > 0.80, f"{what}: only {frac_exact:.3f} of elements bit-equal"


def make(M, N, K, dtype, dev, seed, wrows=None):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(M, K, generator=g) * 0.5).to(dtype).to(dev)
    w = (torch.randn(wrows or N, K, generator=g) * 0.05).to(dtype).to(dev)
    return x, w


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M", [1, 2, 3, 4])
def test_gemv_matches_reference(lib, cuda_dev, dtype, M):
    for (N, K) in [(256, 256), (1001, 704), (4096, 4096)]:
        x, w = make(M, N, K, dtype, cuda_dev, seed=N + K + M)
        out = run_linear(lib, x, w, M, N, K, dtype, _lib.ALGO_GEMV)
        assert_close_ulp(out, ref_linear(x, w, dtype), dtype, f"gemv M={M} N={N} K={K}")


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M", [5, 16, 32, 33, 64, 100, 128, 200, 256, 300, 1000])
def test_tcgen05_plain(lib, cuda_dev, dtype, M):
    ws = torch.zeros(64 << 20, dtype=torch.uint8, device=cuda_dev)
    for (N, K) in [(128, 64), (256, 256), (1001, 704), (768, 3072)]:
        x, w = make(M, N, K, dtype, cuda_dev, seed=N + K + M)
        out = run_linear(lib, x, w, M, N, K, dtype, _lib.ALGO_TC, ws=ws)
        assert_close_ulp(out, ref_linear(x, w, dtype), dtype, f"tc M={M} N={N} K={K}")


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("splits", [1, 2, 3, 7, 8, 12])
def test_tcgen05_split_k_is_deterministic_and_correct(lib, cuda_dev, splits, mode):
    """mode 0: cluster + DSMEM reduction (splits <= 8), mode 1: global fp32 workspace."""
    dtype = torch.float16
    M, N, K = 32, 512, 4096
    ws = torch.zeros(64 << 20, dtype=torch.uint8, device=cuda_dev)
    x, w = make(M, N, K, dtype, cuda_dev, seed=7)
    lib.rd_linear_force_splits(splits)
    lib.rd_linear_splitk_mode(mode)
    try:
        a = run_linear(lib, x, w, M, N, K, dtype, _lib.ALGO_TC, ws=ws)
        b = run_linear(lib, x, w, M, N, K, dtype, _lib.ALGO_TC, ws=ws)
    finally:
        lib.rd_linear_force_splits(0)
        lib.rd_linear_splitk_mode(0)
    assert torch.equal(a, b), "split-K result changed between two launches"
    assert_close_ulp(a, ref_linear(x, w, dtype), dtype, f"tc split={splits}")


@pytest.mark.parametrize("algo", [_lib.ALGO_TC, _lib.ALGO_SIMT, _lib.ALGO_GEMV])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_epilogues(lib, cuda_dev, algo, dtype):
    M = 4 if algo == _lib.ALGO_GEMV else 48
    N, K = 384, 512
    ws = torch.zeros(64 << 20, dtype=torch.uint8, device=cuda_dev)
    g = torch.Generator().manual_seed(5)
    x, w = make(M, N, K, dtype, cuda_dev, seed=11)
    bias = (torch.randn(N, generator=g) * 0.1).to(cuda_dev)
    res = (torch.randn(M, N, generator=g)).to(dtype).to(cuda_dev)
    # bias + relu / gelu
    for act in (_lib.ACT_RELU, _lib.ACT_GELU, _lib.ACT_NONE):
        out = run_linear(lib, x, w, M, N, K, dtype, algo, bias=bias, act=act, ws=ws)
        assert_close_ulp(out, ref_linear(x, w, dtype, bias=bias, act=act), dtype, f"algo{algo} bias act{act}")
    # residual, both rounding modes
    for mode in (1, 2):
        out = run_linear(lib, x, w, M, N, K, dtype, algo, bias=bias if mode == 2 else None, residual=res, res_mode=mode, ws=ws)
        ref = ref_linear(x, w, dtype, bias=bias if mode == 2 else None, residual=res, res_mode=mode)
        assert_close_ulp(out, ref, dtype, f"algo{algo} residual mode{mode}", mag=ref.float().abs() + res.float().abs())
    # LoRA side product
    lt = (torch.randn(M, 16, generator=g) * 0.3).to(dtype).to(cuda_dev)
    lb = (torch.randn(N, 16, generator=g) * 0.05).to(dtype).to(cuda_dev)
    out = run_linear(lib, x, w, M, N, K, dtype, algo, lora_t=lt, lora_b=lb, lora_scale=2.0, ws=ws)
    ref = ref_linear(x, w, dtype, lora_t=lt, lora_b=lb, lora_scale=2.0)
    assert_close_ulp(out, ref, dtype, f"algo{algo} lora", mag=ref.float().abs() + 2.0 * (lt.float() @ lb.float().t()).abs())
    # SwiGLU (gate rows then up rows)
    x2, w2 = make(M, N, K, dtype, cuda_dev, seed=13, wrows=2 * N)
    out = run_linear(lib, x2, w2, M, N, K, dtype, algo, act=_lib.ACT_SWIGLU, ws=ws)
    ref = ref_linear(x2, w2, dtype, act=_lib.ACT_SWIGLU, N=N)
    assert torch.isfinite(out.float()).all()
    eps = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    tol = 4 * eps * ref.float().abs().clamp_min(1e-2)
    assert ((out.float() - ref.float()).abs() <= tol).all(), f"algo{algo} swiglu"


def test_vicuna_decode_shapes_tc_vs_gemv_vs_simt(lib, cuda_dev):
    """The five GEMM shapes of a Vicuna-7B decode step at batch 32 / 4: tcgen05, GEMV and SIMT agree with each other."""
    dtype = torch.float16
    ws = torch.zeros(256 << 20, dtype=torch.uint8, device=cuda_dev)
    for (N, K, act) in [(12288, 4096, 0), (4096, 4096, 0), (11008, 4096, _lib.ACT_SWIGLU), (4096, 11008, 0), (32001, 4096, 0)]:
        wrows = 2 * N if act else N
        x, w = make(32, N, K, dtype, cuda_dev, seed=N, wrows=wrows)
        tc = run_linear(lib, x, w, 32, N, K, dtype, _lib.ALGO_TC, act=act, ws=ws)
        simt = run_linear(lib, x, w, 32, N, K, dtype, _lib.ALGO_SIMT, act=act)
        gemv = run_linear(lib, x[:4].contiguous(), w, 4, N, K, dtype, _lib.ALGO_GEMV, act=act)
        eps = 2.0 ** -10
        for name, a, b in (("tc-simt", tc, simt), ("gemv-simt", gemv, simt[:4])):
            assert torch.isfinite(a.float()).all() and torch.isfinite(b.float()).all(), f"{name} N={N} K={K}: unwritten output"
            # SwiGLU multiplies two rounded values: an ulp of the larger factor can dominate a small product
            floor = 5e-2 if act else 1e-2
            err = (a.float() - b.float()).abs() / b.float().abs().clamp_min(floor)
            assert err.max().item() <= 4 * eps, f"{name} N={N} K={K}: {err.max().item():.3e}"


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K,act,res", [(32, 4096, 4096, 0, True), (32, 1024, 4096, _lib.ACT_SWIGLU, False), (32, 512, 11008, 0, True),
                                           (7, 12304, 1024, 0, False), (1, 4096, 4096, 0, True), (17, 2048, 704, _lib.ACT_SWIGLU, False),
                                           (32, 32001, 512, 0, False)])
def test_tmem_staged_decode_tiles_bit_identical(cuda_dev, lib, dtype, M, N, K, act, res):
    """Decode tiles with the weight k-blocks parked in tensor memory (tcgen05.st by the epilogue warps, tcgen05.mma with the A
    operand from TMEM) against the both-operands-from-shared-memory kernel: same products, same accumulation order, same
    split-K reduction, so the outputs must be bit-identical - and both must match the fp32 reference."""
    g = torch.Generator().manual_seed(M * 7 + N + K)
    rows = 2 * N if act == _lib.ACT_SWIGLU else N
    w = (torch.randn(rows, K, generator=g) * 0.05).to(dtype).to(cuda_dev)
    x = (torch.randn(M, K, generator=g) * 0.5).to(dtype).to(cuda_dev)
    residual = (torch.randn(M, N, generator=g) * 0.5).to(dtype).to(cuda_dev) if res else None
    ws = torch.zeros(int(lib.rd_linear_workspace_bytes(M, N, K)) + 256, dtype=torch.uint8, device=cuda_dev)
    outs = {}
    for ts in (0, 1):
        lib.rd_linear_tmem_staging(ts)
        outs[ts] = [run_linear(lib, x, w, M, N, K, dtype, _lib.ALGO_TC, act=act, residual=residual, ws=ws) for _ in range(2)]
    lib.rd_linear_tmem_staging(0)
    assert torch.equal(outs[1][0], outs[1][1]), "TMEM-staged kernel is not deterministic run to run"
    assert torch.equal(outs[0][0], outs[1][0]), f"max diff {(outs[0][0].float() - outs[1][0].float()).abs().max().item():.4g}"
    ref = ref_linear(x, w, dtype, act=act, residual=residual, N=N)
    ulp = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    scale = ref.float().abs().max().item()
    assert (outs[1][0].float() - ref.float()).abs().max().item() <= 2 * ulp * scale


WIDE_CASES = [
    # M, N, K, act, bias, res_mode (0 = no residual)
    (2048, 4096, 4096, 0, False, 1),            # prefill o_proj
    (2048, 1536, 1024, _lib.ACT_SWIGLU, False, 0),
    (1000, 1000, 704, 0, False, 0),             # ragged M / N / K tails
    (777, 328, 72, _lib.ACT_RELU, True, 2),     # conv-like: bias + fp32 residual + ReLU, K barely over one k-block
    (3136, 256, 64, 0, True, 0),                # one k-block per tile (layer-1 1x1 conv)
    (1290, 776, 256, _lib.ACT_GELU, True, 0),
    (515, 520, 1152, _lib.ACT_SWIGLU, False, 0),
    (300, 4096, 512, 0, False, 1),
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("case", WIDE_CASES)
@pytest.mark.parametrize("nt,stages", [(0, 0), (128, 2), (160, 3), (240, 4)])
def test_persistent_wide_kernel_bit_identical(cuda_dev, lib, dtype, case, nt, stages):
    """linear_wide.cu (persistent CTAs, double-buffered TMEM accumulators, TMA residual load / store, runtime token-tile
    width) against the one-tile-per-CTA kernel: same products, same k-block order, same epilogue arithmetic, so bit-identical
    - and both within an ulp of the fp32 reference.  min_tiles = 1 forces the persistent kernel onto the small shapes too."""
    M, N, K, act, has_bias, res_mode = case
    g = torch.Generator().manual_seed(M + 3 * N + 5 * K + nt)
    rows = 2 * N if act == _lib.ACT_SWIGLU else N
    w = (torch.randn(rows, K, generator=g) * 0.05).to(dtype).to(cuda_dev)
    x = (torch.randn(M, K, generator=g) * 0.5).to(dtype).to(cuda_dev)
    bias = (torch.randn(N, generator=g) * 0.1).to(cuda_dev) if has_bias else None
    residual = (torch.randn(M, N, generator=g) * 0.5).to(dtype).to(cuda_dev) if res_mode else None
    ws = torch.zeros(int(lib.rd_linear_workspace_bytes(M, N, K)) + 256, dtype=torch.uint8, device=cuda_dev)
    kw = dict(act=act, bias=bias, residual=residual, res_mode=res_mode or 1, ws=ws)
    outs = {}
    try:
        lib.rd_linear_wide_min_tiles(1)
        lib.rd_linear_wide_force_nt(nt)
        lib.rd_linear_wide_force_stages(stages)
        for on in (0, 1, 2):               # 0: one tile per CTA, 1: persistent CTA pairs (cta_group::2), 2: persistent single CTAs
            lib.rd_linear_wide_persistent(1 if on else 0)
            lib.rd_linear_wide_pair(1 if on == 1 else 0)
            lib.rd_linear_force_splits(0 if on else 1)      # the one-tile-per-CTA kernel without split-K: one summation order
            outs[on] = [run_linear(lib, x, w, M, N, K, dtype, _lib.ALGO_TC, **kw) for _ in range(2)]
    finally:
        lib.rd_linear_force_splits(0)
        lib.rd_linear_wide_persistent(1)
        lib.rd_linear_wide_pair(1)
        lib.rd_linear_wide_min_tiles(1)
        lib.rd_linear_wide_force_nt(0)
        lib.rd_linear_wide_force_stages(0)
    assert torch.isfinite(outs[1][0].float()).all(), "persistent kernel left outputs unwritten"
    assert torch.equal(outs[1][0], outs[1][1]), "persistent kernel is not deterministic run to run"
    assert torch.equal(outs[0][0], outs[1][0]), f"pair: max diff {(outs[0][0].float() - outs[1][0].float()).abs().max().item():.4g}"
    assert torch.equal(outs[0][0], outs[2][0]), f"single: max diff {(outs[0][0].float() - outs[2][0].float()).abs().max().item():.4g}"
    assert torch.equal(outs[2][0], outs[2][1]), "persistent single-CTA kernel is not deterministic run to run"
    ref = ref_linear(x, w, dtype, bias=bias, act=act, residual=residual, res_mode=res_mode or 1, N=N)
    ulp = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    scale = ref.float().abs().max().item()
    assert (outs[1][0].float() - ref.float()).abs().max().item() <= 2 * ulp * scale


def test_persistent_wide_kernel_many_tiles_per_cta(cuda_dev, lib):
    """More than two rounds of tiles per CTA (accumulator-buffer and chunk-buffer phases wrap several times), on the prefill
    gate|up shape, against the fp32 reference and the one-tile-per-CTA kernel."""
    dtype = torch.bfloat16
    M, N, K = 2048, 11008, 512
    g = torch.Generator().manual_seed(11)
    w = (torch.randn(2 * N, K, generator=g) * 0.05).to(dtype).to(cuda_dev)
    x = (torch.randn(M, K, generator=g) * 0.5).to(dtype).to(cuda_dev)
    a = run_linear(lib, x, w, M, N, K, dtype, _lib.ALGO_TC, act=_lib.ACT_SWIGLU)
    lib.rd_linear_wide_persistent(0)
    try:
        b = run_linear(lib, x, w, M, N, K, dtype, _lib.ALGO_TC, act=_lib.ACT_SWIGLU)
    finally:
        lib.rd_linear_wide_persistent(1)
    assert torch.equal(a, b)
    ref = ref_linear(x, w, dtype, act=_lib.ACT_SWIGLU, N=N)
    assert (a.float() - ref.float()).abs().max().item() <= 2 * 2.0 ** -7 * ref.float().abs().max().item()


CONV_CASES = [
    # B, H, W, C, Cout, ks, stride, pad, act, residual
    (2, 56, 56, 64, 64, 3, 1, 1, _lib.ACT_RELU, False),      # layer1 conv2 (one half-empty weight tile)
    (3, 28, 20, 128, 128, 3, 2, 1, _lib.ACT_RELU, False),    # stride 2, H != W, tokens not a multiple of the tile
    (2, 28, 28, 256, 256, 3, 1, 1, _lib.ACT_RELU, False),    # two weight tiles, K = 2304: CTA pairs
    (1, 14, 14, 512, 512, 3, 1, 1, _lib.ACT_RELU, False),    # 196 output pixels: one ragged tile
    (2, 56, 56, 256, 512, 1, 2, 0, 0, False),                # stride-2 1x1 downsample
    (5, 9, 11, 64, 136, 3, 1, 1, 0, True),                   # odd sizes, Cout not a multiple of 128, fp32 residual
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("case", CONV_CASES)
def test_implicit_gemm_convolution_bit_identical_to_im2col_path(cuda_dev, lib, dtype, case):
    """rd_conv_nhwc_implicit (im2col-mode TMA loads inside the persistent GEMM's producer) against rd_im2col_nhwc + rd_linear
    (same products, same k order: bit-identical) and against torch.nn.functional.conv2d in fp32."""
    B, H, W, Cin, Cout, ks, stride, pad, act, has_res = case
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + Cin + ks)
    x = (torch.randn(B, H, W, Cin, generator=g) * 0.5).to(dtype).to(cuda_dev)
    w = (torch.randn(Cout, ks, ks, Cin, generator=g) * 0.05).to(dtype).to(cuda_dev)      # [Cout, kh, kw, c] = the im2col column order
    bias = (torch.randn(Cout, generator=g) * 0.1).to(cuda_dev)
    OH, OW = (H + 2 * pad - ks) // stride + 1, (W + 2 * pad - ks) // stride + 1
    M, K = B * OH * OW, ks * ks * Cin
    residual = (torch.randn(M, Cout, generator=g) * 0.5).to(dtype).to(cuda_dev) if has_res else None
    e = _lib.Epilogue()
    e.bias_dev = _lib.ptr(bias)
    e.residual_dev = _lib.ptr(residual)
    e.ld_res = Cout
    e.res_mode = 2
    e.act = act
    out_i = torch.full((M, Cout), float("nan"), device=cuda_dev, dtype=dtype)
    r = lib.rd_conv_nhwc_implicit(_lib.ptr(x), _lib.ptr(w), _lib.ptr(out_i), Cout, B, H, W, Cin, Cout, ks, stride, pad, C.byref(e),
                                  _lib.dtype_code(dtype), _lib.current_stream())
    assert r == 1, f"implicit path declined this shape (r={r}): {lib.rd_last_error()}"
    torch.cuda.synchronize()
    col = torch.empty(M, K, device=cuda_dev, dtype=dtype)
    _lib.check(lib.rd_im2col_nhwc(_lib.ptr(x), _lib.ptr(col), B, H, W, Cin, ks, stride, pad, _lib.dtype_code(dtype), _lib.current_stream()), "im2col")
    lib.rd_linear_wide_min_tiles(1)          # the explicit GEMM through the same persistent kernel (no split-K): same summation order
    try:
        out_e = run_linear(lib, col, w.reshape(Cout, K), M, Cout, K, dtype, _lib.ALGO_TC, bias=bias, act=act, residual=residual, res_mode=2)
    finally:
        lib.rd_linear_wide_min_tiles(1)
    assert torch.isfinite(out_i.float()).all(), "implicit conv left outputs unwritten"
    assert torch.equal(out_i, out_e), f"max diff {(out_i.float() - out_e.float()).abs().max().item():.4g}"
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, stride=stride, padding=pad)
    ref = ref.permute(0, 2, 3, 1).reshape(M, Cout)
    if has_res:
        ref = ref + residual.float()
    if act == _lib.ACT_RELU:
        ref = torch.relu(ref)
    ulp = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    assert (out_i.float() - ref).abs().max().item() <= 2 * ulp * ref.abs().max().item() + 1e-3
