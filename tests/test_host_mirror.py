"""The C++ host mirror (simplellminference_b200/host): builds everywhere; its C++ test driver runs on the GPU box."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "simplellminference_b200", "lib", "libsllm_host.so")
DRIVER = os.path.join(ROOT, "tests", "cpp", "_build", "test_host_mirror")


def _build(only_if_missing=False):
    """only_if_missing: on the GPU box the artefacts arrive prebuilt (build() ran where the snapshot was taken); if the copy did not
    keep modification times, make would recompile every CUDA file serially there (minutes of nvcc) for nothing."""
    lib = os.path.join(ROOT, "simplellminference_b200", "lib", "libsllm_b200.so")
    port = os.path.join(ROOT, "oracle", "_build", "liboracle_port.so")
    if only_if_missing and all(os.path.exists(f) for f in (lib, HOST_SO, port, DRIVER)):
        return
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "simplellminference_b200", "csrc")], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "simplellminference_b200", "host")], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")], check=True)


def test_host_mirror_builds_and_exports_reference_api():
    _build()
    out = subprocess.run(["nm", "-DC", "--defined-only", HOST_SO], capture_output=True, text=True, check=True).stdout
    for sym in ("kernel::add_kernel_cuda", "kernel::emb_kernel_cuda", "kernel::matmul_kernel_cuda", "kernel::mha_kernel_cuda",
                "kernel::rmsnorm_kernel_cuda", "kernel::rope_cache_cal_cuda", "kernel::rope_kernel_cuda", "kernel::swiglu_kernel_cuda",
                "op::VecAddLayer::forward()", "op::EmbeddingLayer::forward()", "op::MatmulLayer::forward()", "op::MultiHeadAttention::forward()",
                "op::RmsNormLayer::forward()", "op::RoPELayer::forward()", "op::SwigluLayer::forward()", "op::argmaxLayer::forward(",
                "mem::Tensor::to_cuda()", "mem::slice_KV_cache(", "mem::CUDADeviceAllocator::allocate(", "model::LlamaModel::forward()",
                "model::LlamaModel::init()", "model::LlamaModel::predict("):
        assert sym in out, f"{sym} missing from libsllm_host.so"
    # the mirror must not carry any CPU kernel (no fallback) and must not link the oracle
    assert not re.search(r"kernel::\w+_cpu", out)
    ldd = subprocess.run(["ldd", HOST_SO], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "libsllm_b200.so" in ldd


def test_compat_headers_cover_the_reference_include_names():
    """Every header name the reference's hot path includes exists as a forwarding header."""
    compat = set(os.listdir(os.path.join(ROOT, "simplellminference_b200", "host", "include", "compat")))
    for name in ("base.h", "alloc.h", "buffer.h", "tensor.h", "layer.h", "add.h", "argmax.h", "embedding.h", "matmul.h", "mha.h",
                 "rmsnorm.h", "rope.h", "swiglu.h", "config.h", "weight_loader.h", "model.h", "add_kernel.cuh", "emb_kernel.cuh",
                 "matmul_kernel.cuh", "mha_kernel.cuh", "rms_kernel.cuh", "rope_kernel.cuh", "swiglu_kernel.cuh"):
        assert name in compat


@pytest.mark.gpu
def test_host_mirror_cpp_driver(tmp_path):
    _build(only_if_missing=True)
    r = subprocess.run([DRIVER], capture_output=True, text=True, cwd=tmp_path, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "PASS" in r.stdout
