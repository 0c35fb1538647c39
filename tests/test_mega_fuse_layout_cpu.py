"""The index arithmetic of the experimental fused down projection of the decode megakernel (SLLM_ENGINE_MEGA_FUSE_DOWN), restated in
numpy: the transposed, stripe-major tile layout written by repack_down_t_kernel (csrc/megakernel.cu), the tile rows a CTA / warp / lane consumes in
the PH_DOWN_T phase, and the split of the gate_up phase that must hand every CTA exactly the units whose columns it then multiplies
(csrc/mega_common.cuh phase_tiles). The sum over all CTAs, warps and lanes must be Wdown . swi, with every input used exactly once.
A restatement cannot prove the CUDA code right (tests/test_zzz_mega_fuse_gpu.py does that on a GPU); it pins the layout contract the
two sides of the kernel were written against."""
import numpy as np
import pytest

JT, WARPS = 4, 16          # csrc/megakernel.cuh kFuseJT, csrc/mega_common.cuh kMegaWarps


def cta_tiles(ntr, cta, ncta):
    return (ntr * cta) // ncta, (ntr * (cta + 1)) // ncta


@pytest.mark.parametrize("d,inter,E,ncta,R_gateup", [(128, 384, 4, 148, 4), (256, 704, 8, 148, 4), (512, 1408, 8, 148, 2),
                                                      (256, 516, 4, 20, 4), (1024, 2824, 8, 148, 4)])
def test_fused_down_layout_and_split(d, inter, E, ncta, R_gateup):
    rng = np.random.default_rng(0)
    W = rng.standard_normal((d, inter)).astype(np.float32)          # Wdown, row-major [d][I] as the loaders leave it
    swi = rng.standard_normal(inter).astype(np.float32)
    KS = d // (32 * E)                                              # stripes of 32 lanes x 16 bytes of outputs
    assert KS in (1, 2, 4, 8, 16) and inter % JT == 0               # mega_fuse_down_ok
    RG = WARPS // KS
    i = np.arange(d * inter)                                        # repack_down_t_kernel: destination element index
    e, lane, jj = i % E, (i // E) % 32, (i // (E * 32)) % JT
    ntr_e, upp = inter // JT, R_gateup // 2
    g, ks = (i // (E * 32 * JT)) % ntr_e, i // (E * 32 * JT * ntr_e)      # stripe-major: [ks][g][jj][lane][e]
    dst = W.reshape(-1)[((ks * 32 + lane) * E + e) * inter + (g * JT + jj)]
    tile = JT * 32 * E                                              # elements of one (tile row, stripe) tile = 2 KB
    rows_per_slot = 2                                               # a 4 KB ring slot = two tile rows of one stripe
    per = JT // upp                                                 # gate_up tile rows per down tile row
    x, used = np.zeros(d), np.zeros(inter, int)
    for cta in range(ncta):
        g0, g1 = cta_tiles(ntr_e, cta, ncta)                        # PH_DOWN_T phase: cta_tiles over I / 4 tile rows
        gd0, gd1 = ((ntr_e * cta) // ncta) * per, ((ntr_e * (cta + 1)) // ncta) * per      # gate_up phase under FUSE: phase_tiles
        u0 = gd0 * upp
        n = max(0, min(inter, gd1 * upp) - u0)
        assert u0 == g0 * JT and n == (g1 - g0) * JT                # the CTA produced exactly the values it multiplies
        xs = swi[u0:u0 + n]                                         # what the gate_up epilogue leaves in shared memory
        used[u0:u0 + n] += 1
        for warp in range(WARPS):
            ks_, rg = warp & (KS - 1), warp // KS
            acc = np.zeros((32, E))
            cnt = g1 - g0
            a, b = g0 + (cnt * rg) // RG, g0 + (cnt * (rg + 1)) // RG     # down_t_rows: a contiguous share of the CTA's tile rows
            ptr, left = (ks_ * ntr_e + a) * tile, (b - a) * tile          # the producer's byte range, in elements
            for jt in range(a, b, rows_per_slot):                         # one ring slot per iteration
                n_el = min(rows_per_slot * tile, left)                    # the last copy may hold a single tile row
                slot = dst[ptr:ptr + n_el]
                ptr, left = ptr + n_el, left - n_el
                nj = min(rows_per_slot, b - jt) * JT
                assert nj * 32 * E == n_el
                for j_ in range(nj):
                    acc += slot[j_ * 32 * E:(j_ + 1) * 32 * E].reshape(32, E).astype(np.float64) * float(xs[(jt - g0) * JT + j_])
            assert left == 0
            x[ks_ * 32 * E:(ks_ + 1) * 32 * E] += acc.reshape(-1)   # red.global.add.v4.f32 into x + (ks * 32 + lane) * E
    assert (used == 1).all()
    assert np.allclose(x, W.astype(np.float64) @ swi.astype(np.float64), atol=1e-9)
