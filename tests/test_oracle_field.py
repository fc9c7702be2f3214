"""CPU: invariants of the oracle's field restatement and the index-equivalence proof the kernels rely on."""
import ctypes

import numpy as np
import torch

from oracle import field as F
from oracle import nerfacc as NA


def test_scalings_fp32_artefact():
    """lse_field.py:58-60 evaluated in fp32 gives 2047 (not 2048) at level 15 -- SURVEY Appendix B.1."""
    s = F.hash_scalings(16, 16, 2048)
    assert s.tolist() == [16, 22, 30, 42, 58, 80, 111, 153, 212, 294, 406, 561, 776, 1072, 1482, 2047]


def test_uint32_wrap_hash_equals_int64_hash():
    """T = 2^k and coordinates >= 0: the low k bits of the wrapping uint32 xor equal the int64 hash modulo T."""
    rng = np.random.RandomState(0)
    xyz = rng.randint(0, 2049, size=(200000, 3)).astype(np.int32)
    xyz[:9] = [[0, 0, 0], [2048, 2048, 2048], [2047, 0, 2048], [1, 1, 1], [2048, 0, 0], [0, 2048, 0], [0, 0, 2048], [1023, 1024, 1025], [16, 16, 16]]
    for log2T in (14, 19, 24):
        T = 1 << log2T
        out = np.zeros(len(xyz), dtype=np.int64)
        lib = ctypes.CDLL(NA.build_c())
        lib.oracle_hash_index(ctypes.c_int64(len(xyz)), xyz.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(T), ctypes.c_int64(0),
                              out.ctypes.data_as(ctypes.c_void_p))
        x = xyz.astype(np.uint32)
        wrap = (x[:, 0] ^ (x[:, 1] * np.uint32(2654435761)) ^ (x[:, 2] * np.uint32(805459861))) & np.uint32(T - 1)
        assert np.array_equal(out, wrap.astype(np.int64))
        # and the torch restatement agrees with the C one
        v = torch.from_numpy(xyz).to(torch.int64) * torch.tensor([1, 2654435761, 805459861])
        assert np.array_equal(((v[:, 0] ^ v[:, 1] ^ v[:, 2]) % T).numpy(), out)


def test_hash_encode_is_interpolation():
    """at lattice points the encoding equals the table row; weights sum to one (constant table -> constant output)."""
    sc = F.hash_scalings()
    T = 2**12
    table = torch.full((16 * T, 2), 0.37)
    x = torch.rand(100, 3)
    assert torch.allclose(F.hash_encode(x, table, sc, T), torch.full((100, 32), 0.37), atol=1e-6)


def test_contraction_and_selector():
    p = torch.tensor([[0.2, -0.3, 0.9], [2.0, 0.0, 0.0], [0.0, -4.0, 1.0], [1e6, 1e6, 1e6]])
    c = F.contract_inf(p)
    assert torch.allclose(c[0], p[0])
    assert torch.allclose(c[1], torch.tensor([1.5, 0.0, 0.0]))
    assert torch.allclose(c[2], torch.tensor([0.0, -1.75, 0.4375]))
    x, sel = F.normalized_positions(p)
    assert sel.tolist() == [True, True, True, True] and float(x.max()) < 1.0


def test_trunc_exp_gradient_clamp():
    x = torch.tensor([-20.0, 0.0, 20.0], requires_grad=True)
    F.trunc_exp(x).sum().backward()
    assert torch.allclose(x.grad, torch.exp(torch.tensor([-15.0, 0.0, 15.0])))


def test_sh_matches_closed_form():
    v = torch.tensor([[0.3, -0.5, 0.8]])
    o = F.sh_deg4(v)[0]
    assert abs(o[0].item() - 0.28209479177387814) < 1e-7
    assert abs(o[6].item() - (0.9461746957575601 * 0.64 - 0.31539156525251999)) < 1e-6
    assert abs(o[10].item() - 2.890611442640554 * 0.3 * -0.5 * 0.8) < 1e-6


def test_sh_constants_against_scipy_real_harmonics():
    """Independent check of the 16 restated SH polynomials / constants (nerfstudio components_from_spherical_harmonics, levels=4): on unit
    vectors they must equal the orthonormal real spherical harmonics built from scipy's complex Y_l^m -- order m = -l..l per
    degree, no Condon-Shortley sign.  (Pins the mathematics, i.e. guards against a mistyped constant shared by oracle and kernel;
    the reference's own convention stays third-party, oracle/__init__.py.)"""
    import numpy as np
    import torch
    from scipy.special import sph_harm_y
    from oracle import field as OF
    g = np.random.RandomState(0)
    v = g.randn(500, 3); v /= np.linalg.norm(v, axis=1, keepdims=True)
    theta, phi = np.arccos(np.clip(v[:, 2], -1, 1)), np.arctan2(v[:, 1], v[:, 0])
    cols = []
    for l in range(4):
        for m in range(-l, l + 1):
            Y = sph_harm_y(l, abs(m), theta, phi)
            if m == 0:
                cols.append(Y.real)
            elif m > 0:
                cols.append(np.sqrt(2.0) * (-1) ** m * Y.real)
            else:
                cols.append(np.sqrt(2.0) * (-1) ** m * Y.imag)
    want = np.stack(cols, axis=1)
    got = OF.sh_deg4(torch.from_numpy(v)).numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-7)


def test_x_corner_pairs_share_a_16_byte_slot():
    """Locality of the hash (the reason coarse-in-x neighbours are cheap; the paired-load / paired-reduction / two-lanes-per-sample kernels
    built on it were measured and removed in round 2, profiles/r2a_ab_pair_*.json): the x term enters the hash unmultiplied, so whenever (idx[0] ^ idx[3]) == 1 every
    (ceil_x, floor_x) corner pair -- (0,3), (1,2), (4,7), (5,6) -- sits on rows j ^ 1 and j of the level's table, i.e. in the two halves of
    the 16-byte slot j >> 1; this happens for about half of the samples at fine levels."""
    import torch
    from oracle import field as OF
    g = torch.Generator().manual_seed(0)
    x = torch.rand(20000, 3, generator=g)
    x[:64] = torch.randint(0, 16, (64, 3), generator=g).float() / 16.0          # lattice points: ceil == floor, no pair
    T = 2**19
    idx, _ = OF.hash_corner_indices(x, OF.hash_scalings(), T)
    idx = idx - (torch.arange(16) * T)[None, :, None]                              # rows inside each level's sub-table
    assert int(idx.min()) >= 0 and int(idx.max()) < T
    pair = (idx[..., 0] ^ idx[..., 3]) == 1
    for c, f in ((0, 3), (1, 2), (4, 7), (5, 6)):
        assert bool(((idx[..., c] ^ idx[..., f]) == 1)[pair].all())
        assert bool(((idx[..., c] >> 1) == (idx[..., f] >> 1))[pair].all())
        assert not bool(((idx[..., c] ^ idx[..., f]) == 1)[~pair].any())
    frac = pair[:, 8:].float().mean().item()
    assert 0.45 < frac < 0.55, frac
    # ... and the same 128-byte line (16 rows) unless floor_x ends in four one bits
    same_line = ((idx[..., 0] >> 4) == (idx[..., 3] >> 4))[64:, 8:].float().mean().item()
    assert 0.92 < same_line < 0.95, same_line
    # reading the slot as 4 floats and picking halves by the row's low bit returns the rows themselves
    table = torch.rand(T, 2, generator=g)
    slots = table.view(T // 2, 4)
    jf, jc = idx[..., 15, 3][pair[:, 15]], idx[..., 15, 0][pair[:, 15]]
    v = slots[jf >> 1]
    odd = (jf & 1).bool()[:, None]
    assert torch.equal(torch.where(odd, v[:, 2:], v[:, :2]), table[jf]) and torch.equal(torch.where(odd, v[:, :2], v[:, 2:]), table[jc])
