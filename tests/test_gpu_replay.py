"""K1 replay gather / add / Philox: bit-exact against the reference fixtures."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from helpers import philox4x32_10, philox_index
from oracle import td3_oracle as O

pytestmark = pytest.mark.gpu


def _stream(rs, n, S, A):
    return [(rs.standard_normal(S), rs.uniform(-1, 1, A), rs.standard_normal(S).astype(np.float32),
             float(rs.standard_normal()), float(rs.uniform() < 0.2)) for _ in range(n)]


def test_add_wraparound_and_sample_match_reference_fixture():
    from td3_b200.my_replay_buffer import ReplayBuffer_featured
    z = np.load(os.path.join(GOLDEN, "replay_sample.npz"))
    rs = np.random.RandomState(int(z["seed"]))
    rb = ReplayBuffer_featured(O.Space(5), O.Space(2), max_size=37)
    for row in _stream(rs, 50, 5, 2):
        rb.add(*row)
    ind = rs.randint(0, 37, size=64)
    assert [rb.ptr, rb.size] == list(z["feat_ptr_size"])
    out = rb.sample(64, indices=ind)
    for k, v in zip(rb.store_np, out):
        assert v.dtype == torch.float32 and v.is_cuda
        assert np.array_equal(v.cpu().numpy(), z["feat_" + k]), k            # bit-exact incl. duplicates
    # device int64 indices take the same path
    out2 = rb.sample(64, indices=torch.as_tensor(ind, device="cuda"))
    assert all(torch.equal(a, b) for a, b in zip(out, out2))


def test_particles_buffer_matches_reference_fixture():
    from td3_b200.my_replay_buffer import ReplayBuffer_particles
    z = np.load(os.path.join(GOLDEN, "replay_sample.npz"))
    rs = np.random.RandomState(int(z["seed"]))
    _stream(rs, 50, 5, 2)
    rs.randint(0, 37, size=64)                                               # replay the fixture's RNG stream
    act = O.Space(2)
    rb = ReplayBuffer_particles((O.Space(3), O.Space(6, 4)), act, max_size=9)
    for _ in range(14):
        rb.add((rs.standard_normal(3), rs.standard_normal((6, 4))), rs.uniform(-1, 1, 2),
               (rs.standard_normal(3), rs.standard_normal((6, 4))), float(rs.standard_normal()), float(rs.uniform() < 0.3))
    ind = rs.randint(0, 9, size=16)
    assert np.array_equal(ind, z["part_indices"])
    assert [rb.ptr, rb.size] == list(z["part_ptr_size"])
    for k, v in zip(rb.store_np, rb.sample(16, indices=ind)):
        assert np.array_equal(v.cpu().numpy(), z["part_" + k]), k


def test_global_numpy_stream_is_the_default_draw():
    from td3_b200.my_replay_buffer import ReplayBuffer_featured
    data = O.synthetic_transitions_featured(200, 4, 2, seed=3)
    rb = ReplayBuffer_featured(O.Space(4), O.Space(2), max_size=256)
    orb = O.ReplayFeatured(O.Space(4), O.Space(2), 256)
    rb.add_batch(**data)
    O.fill_featured(orb, data)
    assert (rb.ptr, rb.size) == (orb.ptr, orb.size)
    np.random.seed(5)
    want = orb.sample(33)
    np.random.seed(5)
    got = rb.sample(33)
    for a, b in zip(got, want):
        assert np.array_equal(a.cpu().numpy(), b.numpy())


def test_sample_empty_raises_value_error():
    from td3_b200.my_replay_buffer import ReplayBuffer_featured
    rb = ReplayBuffer_featured(O.Space(3), O.Space(1), max_size=8)
    with pytest.raises(ValueError):
        rb.sample(4)


def test_large_rows_and_large_buffer_roundtrip():
    """1M-row buffer (BASELINE cfg 2 size): gather == torch index_select on the same device rows."""
    from td3_b200.my_replay_buffer import ReplayBuffer_featured
    rb = ReplayBuffer_featured(O.Space(17), O.Space(6), max_size=1_000_000)
    g = torch.Generator(device="cuda").manual_seed(0)
    rb._rows.copy_(torch.randn(rb._rows.shape, device="cuda", generator=g))
    rb.size, rb.ptr = rb.max_size, 0
    idx = torch.randint(0, rb.max_size, (4096,), device="cuda", generator=g)
    s, a, s2, r, nd = rb.sample(4096, indices=idx)
    rows = rb._rows[idx]
    assert torch.equal(s, rows[:, 0:17]) and torch.equal(a, rows[:, 17:23]) and torch.equal(s2, rows[:, 23:40])
    assert torch.equal(r, rows[:, 40:41]) and torch.equal(nd, rows[:, 41:42])


def test_philox_known_answer_and_range():
    import ctypes as C
    from td3_b200 import _lib
    # published Philox4x32-10 known-answer vectors (Random123 kat_vectors)
    assert philox4x32_10([0, 0, 0, 0], 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox4x32_10([0xffffffff] * 4, 0xffffffffffffffff) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    lib = _lib.require_cuda()
    n, size, seed, step = 4096, 1_000_000, 0x1234567890ABCDEF, 77
    idx = torch.empty(n, dtype=torch.int64, device="cuda")
    _lib.check(lib.rb_philox_indices(C.c_void_p(idx.data_ptr()), n, size, seed, 0, step, _lib.stream_ptr()))
    got = idx.cpu().numpy()
    want = np.array([philox_index(seed, 0, step, e, size) for e in range(n)])
    assert np.array_equal(got, want)                                         # device == host Philox, bit-exact
    assert got.min() >= 0 and got.max() < size
    # uniformity: chi-square over 16 bins, 64k draws (df=15, 99.9th percentile = 37.7)
    n = 65536
    idx = torch.empty(n, dtype=torch.int64, device="cuda")
    _lib.check(lib.rb_philox_indices(C.c_void_p(idx.data_ptr()), n, size, seed, 0, step + 1, _lib.stream_ptr()))
    counts = np.bincount((idx.cpu().numpy() * 16 // size), minlength=16)
    chi2 = float(((counts - n / 16) ** 2 / (n / 16)).sum())
    assert chi2 < 37.7, chi2


def test_buffer_save_load_reference_format(tmp_path):
    """Files are float64 np.save payloads named <field>.pkl + pickled ptr/size (my_replay_buffer.py:91-107)."""
    import pickle
    from td3_b200.my_replay_buffer import ReplayBuffer_featured
    data = O.synthetic_transitions_featured(40, 4, 2, seed=1)
    rb = ReplayBuffer_featured(O.Space(4), O.Space(2), max_size=64)
    rb.add_batch(**data)
    rb.save(str(tmp_path))
    with open(tmp_path / "state.pkl", "rb") as f:
        arr = np.load(f)
    assert arr.dtype == np.float64 and arr.shape == (64, 4)
    assert np.array_equal(arr[:40], data["state"].astype(np.float32).astype(np.float64))
    with open(tmp_path / "size.pkl", "rb") as f:
        assert pickle.load(f) == 40
    rb2 = ReplayBuffer_featured(O.Space(4), O.Space(2), max_size=64, load_folder=str(tmp_path))
    assert (rb2.ptr, rb2.size) == (40, 40)
    assert torch.equal(rb2._rows, rb._rows)


def test_device_rng_statistics():
    """T5: what the on-device Philox stream feeds an update -- replay indices uniform over [0, size)
    (chi-square over 64 bins, range), smoothing noise = clip(N(0, policy_noise^2), +-noise_clip) with the mean, standard
    deviation and clip fraction of that distribution (TD3_featured.py:131-133)."""
    import math
    from helpers import make_featured
    rows, B, n_upd = 4096, 4096, 8
    _, _, ours, rb = make_featured(rows=rows, actor_widths=(64, 64), q_widths=(64, 64), policy_noise=0.2, noise_clip=0.5)
    idx, eps = [], []
    for _ in range(n_upd):
        ours.train(rb, B)
        torch.cuda.synchronize()
        d = ours.debug_tensors()
        idx.append(d["indices"].cpu().numpy().reshape(-1).copy())
        eps.append(d["eps"].cpu().numpy().reshape(-1).copy())
    idx, eps = np.concatenate(idx), np.concatenate(eps).astype(np.float64)
    assert idx.min() >= 0 and idx.max() < rows
    assert len(np.unique(idx[:B])) < B and not np.array_equal(idx[:B], idx[B:2 * B])     # with replacement, fresh per update
    counts = np.bincount(idx * 64 // rows, minlength=64)
    chi2 = float(((counts - len(idx) / 64) ** 2 / (len(idx) / 64)).sum())
    assert chi2 < 63 + 6 * math.sqrt(2 * 63), chi2                                       # 6 sigma of chi-square(63)
    n = len(eps)
    assert abs(eps).max() <= 0.5 + 1e-7
    clip_frac = float((np.abs(eps) >= 0.5 - 1e-7).mean())
    want_clip = math.erfc(2.5 / math.sqrt(2))                                            # P(|z| > 2.5) = 1.24 %
    assert abs(clip_frac - want_clip) < 6 * math.sqrt(want_clip / n), (clip_frac, want_clip)
    assert abs(eps.mean()) < 6 * 0.2 / math.sqrt(n)
    # variance of a normal clipped at c = 2.5 sigma: sigma^2 (1 - 2 c phi(c) - 2 (1 - c^2) Q(c)) with Q the upper tail
    c = 2.5
    phi, Q = math.exp(-c * c / 2) / math.sqrt(2 * math.pi), 0.5 * math.erfc(c / math.sqrt(2))
    want_std = 0.2 * math.sqrt(1 - 2 * c * phi - 2 * (1 - c * c) * Q)
    assert abs(eps.std() - want_std) < 0.01 * want_std, (eps.std(), want_std)
