"""GPU parity tests, op level: every launcher of include/sllm_b200.h against the CPU oracle (the C restatement,
which tests/test_oracle_cpu.py pins bit-for-bit to the reference) and against the golden fixtures produced by
the reference itself. Tolerances: the GPU sums in a different order than the oracle's serial loops, nothing
else differs (fp32, IEEE divide/sqrt, accurate expf), so |err| <= 1e-5 * (1 + sum|terms|) style bounds hold;
integer/byte work (embedding gather, argmax, weight generation, bf16/int8 conversion) must be bit-exact."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from simplellminference_b200 import kernels as K
from simplellminference_b200 import _lib
from simplellminference_b200.config import F32, BF16, INT8, PRESETS

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def close(got, want, rtol=2e-5, atol=2e-5):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol)


def test_golden_ops(golden_ops):
    """The reference's own outputs (tests/golden/ops_ref.npz) on the seeded inputs."""
    i, g = mg.op_inputs(), golden_ops
    close(K.rmsnorm(dev(i["x"]), dev(i["w"]), i["eps"]), g["rmsnorm"])
    close(K.matmul(dev(i["x"]), dev(i["W"]), *i["W"].shape), g["matmul"], atol=1e-4)
    assert np.array_equal(K.add(dev(i["a"]), dev(i["b"])).cpu().numpy(), g["add"])          # one fp32 add: exact
    close(K.swiglu(dev(i["up"]), dev(i["gate"])), g["swiglu"], rtol=1e-6, atol=1e-6)
    assert np.array_equal(K.embedding(i["token"], dev(i["table"])).cpu().numpy(), g["embedding"])
    s, c = K.rope_tables(i["hd"], i["S"], i["theta"])
    assert np.array_equal(s.cpu().numpy(), g["sin"]) and np.array_equal(c.cpu().numpy(), g["cos"])  # host libm: exact
    q, k = K.rope(dev(i["q"]), dev(i["k"]), i["pos"], s, c, i["hd"])
    close(q, g["rope_q"], rtol=1e-6, atol=1e-6)
    close(k, g["rope_k"], rtol=1e-6, atol=1e-6)
    close(K.mha(dev(i["q"]), dev(i["kc"]), dev(i["vc"]), i["layer"], i["pos"], i["hd"], i["H"], i["KVH"]), g["mha"], atol=1e-5)
    assert int(K.argmax(dev(i["logits"])).item()) == int(g["argmax"]) == 411


@pytest.mark.parametrize("rows,cols", [(1, 48), (7, 288), (288, 288), (77, 768), (2048, 2048), (33, 5632), (4096, 11008)])
@pytest.mark.parametrize("wd", [F32, BF16, INT8])
def test_gemv_vs_oracle(port, rows, cols, wd):
    if wd == INT8 and cols % 64:
        pytest.skip("int8 needs cols % group == 0")
    if wd == BF16 and cols % 8:
        pytest.skip("bf16 needs cols % 8 == 0")
    rng = np.random.default_rng(rows * 131 + cols)
    x = rng.standard_normal(cols).astype(np.float32)
    W = (rng.standard_normal((rows, cols)) * (4.0 / np.sqrt(cols))).astype(np.float32)
    Wd, sc = K.convert_weights(dev(W), wd, 64)
    # the oracle sees exactly the values the GPU stores
    if wd == F32:
        Weff = W
    elif wd == BF16:
        Weff = Wd.float().cpu().numpy()
    else:
        Weff = (Wd.float() * sc.repeat_interleave(64, dim=1)).cpu().numpy()
    want = port.matmul(x, np.ascontiguousarray(Weff), 0.125)
    got = K.matmul(dev(x), Wd, rows, cols, wd, sc, 64, scale=0.125)
    bound = 1e-6 * (np.abs(Weff) @ np.abs(x)) * 0.125 * np.sqrt(cols) + 1e-6
    err = np.abs(got.cpu().numpy() - want)
    assert np.all(err <= bound), (err.max(), bound.min())


def test_convert_weights_bit_exact(port):
    """bf16 = RNE, int8 = rint(w/(amax/127)): identical to the oracle's CPU arithmetic."""
    shape = mg.SHAPES["tiny_gqa"]
    raw = port.fill_segment(shape, 2, 0, 128 * 128)
    Wd, _ = K.convert_weights(dev(raw.reshape(128, 128)), BF16)
    assert np.array_equal(Wd.float().cpu().numpy().ravel(), port.fill_segment(shape, 2, 0, 128 * 128, wdtype=1))
    q, sc = K.convert_weights(dev(raw.reshape(128, 128)), INT8, 64)
    q_ref, sc_ref = port.fill_segment_int8(shape, 2, 0, 128 * 128, group=64)
    assert np.array_equal(q.cpu().numpy().ravel(), q_ref) and np.array_equal(sc.cpu().numpy().ravel(), sc_ref)


@pytest.mark.parametrize("wd", [F32, BF16, INT8])
def test_synth_fill_bit_exact(port, wd):
    """Device generator == CPU generator, including a tensor-parallel column slice."""
    ms = PRESETS["tiny_gqa"]
    shape = mg.SHAPES["tiny_gqa"]
    lib = _lib.load()
    cshape = _lib.Shape(ms.vocab, ms.head_dim, ms.hidden, ms.kv_hidden, ms.inter, ms.max_len, ms.layers, ms.heads, ms.kv_heads, ms.eps, ms.theta)
    tdt = {F32: torch.float32, BF16: torch.bfloat16, INT8: torch.int8}[wd]
    # segment 8 = down [L][d][I]: rows 128..256 (layer 1), columns 192..384 (rank 1 of 2)
    d, I = ms.hidden, ms.inter
    dst = torch.empty(d, I // 2, dtype=tdt, device="cuda")
    sc = torch.empty(d, I // 2 // 64, dtype=torch.float32, device="cuda") if wd == INT8 else None
    import ctypes as C
    _lib.check(lib.sllm_synth_fill(C.byref(cshape), 1234, 8, d, d, I, I // 2, I // 2, dst.data_ptr(), wd,
                                   sc.data_ptr() if sc is not None else None, 64, torch.cuda.current_stream().cuda_stream))
    full = port.fill_segment(shape, 8, d * I, d * I, seed=1234, wdtype=wd, group=64).reshape(d, I)[:, I // 2:]
    got = dst.float() if wd != INT8 else dst.float() * sc.repeat_interleave(64, dim=1)
    assert np.array_equal(got.cpu().numpy(), full)


@pytest.mark.parametrize("hd,H,KVH,S,pos", [(48, 6, 6, 64, 0), (48, 6, 6, 64, 63), (64, 12, 12, 1024, 255), (64, 32, 4, 2048, 2047),
                                             (128, 32, 32, 4096, 511), (128, 32, 8, 1024, 777), (128, 8, 1, 512, 130), (32, 4, 2, 48, 46)])
@pytest.mark.parametrize("kvd", [F32, BF16])
def test_mha_vs_oracle(port, hd, H, KVH, S, pos, kvd):
    rng = np.random.default_rng(hd + H + pos)
    L, layer = 2, 1
    kv = KVH * hd
    q = rng.standard_normal(H * hd).astype(np.float32)
    kc = rng.standard_normal((L, S, kv)).astype(np.float32)
    vc = rng.standard_normal((L, S, kv)).astype(np.float32)
    if kvd == BF16:
        kcd, vcd = dev(kc).bfloat16(), dev(vc).bfloat16()
        kc, vc = kcd.float().cpu().numpy(), vcd.float().cpu().numpy()
    else:
        kcd, vcd = dev(kc), dev(vc)
    want = port.mha(q, kc, vc, layer, pos, hd, H, KVH)
    ws = K.mha_workspace(H, hd, S)
    for _ in range(2):  # twice: the workspace must come back zeroed
        got = K.mha(dev(q), kcd, vcd, layer, pos, hd, H, KVH, workspace=ws, kv_dtype=kvd)
        close(got, want, rtol=1e-4, atol=2e-5)
    # device-side position gives the same answer
    got = K.mha(dev(q), kcd, vcd, layer, torch.tensor([pos], dtype=torch.int32, device="cuda"), hd, H, KVH, workspace=ws, kv_dtype=kvd)
    close(got, want, rtol=1e-4, atol=2e-5)


def test_small_ops_vs_oracle(port):
    rng = np.random.default_rng(5)
    for n in (1, 7, 288, 4096, 11008):
        a, b = rng.standard_normal(n).astype(np.float32), (8 * rng.standard_normal(n)).astype(np.float32)
        if n % 4 == 0:
            assert np.array_equal(K.add(dev(a), dev(b)).cpu().numpy(), port.add(a, b))
        close(K.swiglu(dev(a), dev(b)), port.swiglu(a, b), rtol=1e-6, atol=1e-7)
        w = (1 + 0.02 * rng.standard_normal(n)).astype(np.float32)
        close(K.rmsnorm(dev(a), dev(w), 1e-5), port.rmsnorm(a, w, 1e-5), rtol=1e-5, atol=1e-6)
    # rope with GQA-shaped k (k_dim < q_dim) and a device-side position
    hd, qd, kd, S = 64, 512, 128, 128
    s, c = K.rope_tables(hd, S, 500000.0)
    ps, pc = port.rope_cache(hd, S, 500000.0)
    assert np.array_equal(s.cpu().numpy(), ps) and np.array_equal(c.cpu().numpy(), pc)
    q, k = rng.standard_normal(qd).astype(np.float32), rng.standard_normal(kd).astype(np.float32)
    wq, wk = port.rope(q, k, 77, ps, pc, hd)
    gq, gk = K.rope(dev(q), dev(k), torch.tensor([77], dtype=torch.int32, device="cuda"), s, c, hd)
    close(gq, wq, rtol=1e-6, atol=1e-6)
    close(gk, wk, rtol=1e-6, atol=1e-6)


def test_argmax_edge_cases():
    for n in (1, 2, 31, 1024, 32000, 128256):
        x = torch.randn(n, device="cuda")
        assert int(K.argmax(x).item()) == int(torch.argmax(x).item())
    x = torch.zeros(5000, device="cuda")
    assert int(K.argmax(x).item()) == 0            # all equal -> first
    x[[4999, 1234, 77]] = 1.0
    assert int(K.argmax(x).item()) == 77           # first maximum wins


def test_embedding_dtypes_and_bounds(port):
    rng = np.random.default_rng(3)
    tab = rng.standard_normal((50, 128)).astype(np.float32)
    for wd in (F32, BF16, INT8):
        Wd, sc = K.convert_weights(dev(tab), wd, 64)
        eff = Wd.float() if wd != INT8 else Wd.float() * sc.repeat_interleave(64, dim=1)
        got = K.embedding(49, Wd, wd, sc, 64)
        assert np.array_equal(got.cpu().numpy(), eff[49].cpu().numpy())
        got = K.embedding(torch.tensor([7], dtype=torch.int32, device="cuda"), Wd, wd, sc, 64)
        assert np.array_equal(got.cpu().numpy(), eff[7].cpu().numpy())
    with pytest.raises(_lib.SllmError, match="outside the vocabulary"):
        K.embedding(50, dev(tab))      # the reference's guard is off by one (emb_kernel.cu:16); ours is not


def test_argument_errors_are_loud():
    x = torch.zeros(100, device="cuda")
    with pytest.raises(_lib.SllmError):
        K.matmul(x, torch.zeros(3, 100, device="cuda", dtype=torch.bfloat16), 3, 100, BF16)   # cols % 8 != 0
    with pytest.raises(_lib.SllmError):
        K.matmul(x[:96], torch.zeros(3, 96, device="cuda", dtype=torch.int8), 3, 96, INT8, None, 64)  # no scales
