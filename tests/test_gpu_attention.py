"""Op-level parity of rd_attention (LlamaAttention core, modeling_llama_imgemb.py:216-234) through the C-ABI: the tcgen05
tensor-core kernel (q_len >= 4, keys <= 256) and the SIMT kernels against the oracle's `attention` + `make_attention_mask`
on the same q / K / V cache / padding mask.  Probabilities are rounded to the storage dtype before P.V in all three, so the
outputs agree to a few storage-dtype ulps (only fp32 summation order differs)."""
import pytest
import torch

from radialog_b200 import _lib
from oracle import radialog_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,nh,q_len,ctx,cmax", [(3, 4, 64, 0, 96), (2, 3, 24, 40, 128), (2, 2, 130, 0, 160), (1, 2, 5, 200, 256), (2, 2, 100, 150, 256)])
def test_attention_tensor_core_and_simt_vs_oracle(cuda_dev, lib, dtype, B, nh, q_len, ctx, cmax):
    g = torch.Generator().manual_seed(B * 1000 + q_len + ctx)
    hd, H = 128, nh * 128
    c_tot = ctx + q_len
    ldq = 3 * H + 16
    qkv = (torch.randn(B * q_len, ldq, generator=g) * 0.7).to(dtype)
    kcache = torch.zeros(B, nh, cmax, hd, dtype=dtype)
    vcache = torch.zeros(B, nh, cmax, hd, dtype=dtype)
    kcache[:, :, :c_tot] = (torch.randn(B, nh, c_tot, hd, generator=g) * 0.7).to(dtype)
    vcache[:, :, :c_tot] = (torch.randn(B, nh, c_tot, hd, generator=g) * 0.7).to(dtype)
    mask = torch.ones(B, cmax, dtype=torch.uint8)
    for b in range(B):                                   # left padding of different lengths (row 0: none)
        mask[b, : (5 * b) % max(1, c_tot - 1)] = 0
    mask[:, c_tot:] = 0
    # oracle: q rows of each sequence against the first c_tot cache rows
    q = qkv[:, :H].view(B, q_len, nh, hd).transpose(1, 2)
    amask = O.make_attention_mask(mask[:, :c_tot].long(), q_len, dtype)
    ref = O.attention(q, kcache[:, :, :c_tot], vcache[:, :, :c_tot], amask, dtype).transpose(1, 2).reshape(B * q_len, H)
    outs = {}
    d = cuda_dev
    qkv_d, k_d, v_d, m_d = qkv.to(d), kcache.to(d), vcache.to(d), mask.to(d)
    ctx_d = torch.tensor([ctx, 0, 0, 0], dtype=torch.int32, device=d)
    for tc in (1, 0):
        lib.rd_attention_set_tensor_core(tc)
        out = torch.zeros(B * q_len, H, dtype=dtype, device=d)
        _lib.check(lib.rd_attention(qkv_d.data_ptr(), ldq, k_d.data_ptr(), v_d.data_ptr(), m_d.data_ptr(), ctx_d.data_ptr(), out.data_ptr(),
                                    B, q_len, nh, hd, cmax, _lib.dtype_code(dtype), torch.cuda.current_stream().cuda_stream), "rd_attention")
        torch.cuda.synchronize()
        outs[tc] = out.cpu().float()
    lib.rd_attention_set_tensor_core(1)
    ulp = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    scale = ref.float().abs().max().item()
    for tc, o in outs.items():
        assert torch.isfinite(o).all()
        err = (o - ref.float()).abs().max().item()
        assert err <= 8 * ulp * scale, f"tensor_core={tc}: max err {err:.4g} vs scale {scale:.4g}"
