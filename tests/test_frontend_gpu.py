"""GPU parity: fused sm_100a frontend kernel (through the C ABI) vs the CPU oracle — bit-exact."""
import numpy as np
import pytest
import torch

from oracle.frontend_oracle import FrontendOracle
from multilingual_kws_b200.synthetic import synthetic_pcm, synthetic_stream

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fe(kws_lib):
    from multilingual_kws_b200.frontend import MicroFrontend
    return MicroFrontend()


@pytest.fixture(scope="module")
def orc():
    return FrontendOracle()


def gpu_u16(fe, pcm_np):
    out = fe.forward(torch.from_numpy(pcm_np).cuda(), raw_u16=True)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_config1_batch32_bit_exact(fe, orc):
    """BASELINE config 1: batch=32 synthetic 1 s clips; 'feature parity <= 1e-4' == integer equality."""
    pcm = synthetic_pcm(32, cfg_id=1)
    want = orc.features_u16(pcm)
    assert np.array_equal(gpu_u16(fe, pcm), want)
    f = fe.forward(torch.from_numpy(pcm).cuda()).cpu().numpy()
    assert f.dtype == np.float32 and f.shape == (32, 49, 40)
    assert np.abs(f - orc.features(pcm)).max() <= 1e-4          # the tolerance north_star states
    assert np.array_equal(f, orc.features(pcm))


def test_adversarial_inputs(fe, orc):
    rng = np.random.default_rng(7)
    adv = rng.integers(-32768, 32768, (16, 16000)).astype(np.int16)
    adv[0] = -32768
    adv[1] = 32767
    adv[2, ::2] = -32768
    adv[2, 1::2] = 32767
    adv[3] = 1
    adv[4] = -1
    adv[5] = 0
    adv[5, 5000] = -32768
    adv[6] = 0
    assert np.array_equal(gpu_u16(fe, adv), orc.features_u16(adv))


def test_batch_1024_and_odd_batches(fe, orc):
    pcm = synthetic_pcm(1024, cfg_id=2)
    assert np.array_equal(gpu_u16(fe, pcm), orc.features_u16(pcm, threads=8))
    for b in (1, 3, 7):
        assert np.array_equal(gpu_u16(fe, pcm[:b]), orc.features_u16(pcm[:b]))


def test_ragged_and_empty(fe, orc):
    pcm = synthetic_pcm(5, cfg_id=4)
    for n in (479, 480, 481, 799, 800, 1234, 15999, 16000, 16001, 20000):
        x = np.ascontiguousarray(np.tile(pcm, (1, 2))[:, :n])
        got = gpu_u16(fe, x)
        want = orc.features_u16(x)
        assert got.shape == want.shape, n
        assert np.array_equal(got, want), n
    assert fe.forward(torch.zeros((0, 16000), dtype=torch.int16, device="cuda")).shape == (0, 49, 40)
    assert fe.forward(torch.zeros((2, 100), dtype=torch.int16, device="cuda")).shape == (2, 0, 40)


def test_long_clips_use_split_path(fe, orc):
    x = synthetic_stream(3 * 200000, cfg_id=6).reshape(3, 200000)     # does not fit the fused kernel's smem
    assert np.array_equal(gpu_u16(fe, x), orc.features_u16(x))


def test_single_vector_and_errors(fe, orc):
    pcm = synthetic_pcm(2)
    one = fe.forward(torch.from_numpy(pcm[1]).cuda()).cpu().numpy()
    assert one.shape == (49, 40) and np.array_equal(one, orc.features(pcm[1:2])[0])
    with pytest.raises(ValueError, match="audio is not a vector"):
        fe.forward(torch.zeros((2, 2, 16000), dtype=torch.int16, device="cuda"))
    with pytest.raises(TypeError):
        fe.forward(torch.zeros((2, 16000), dtype=torch.float32, device="cuda"))


@pytest.mark.parametrize("kw_o,kw_g", [
    (dict(enable_pcan=0), dict(enable_pcan=False)),
    (dict(enable_log=0), dict(enable_log=False)),
    (dict(num_channels=32, window_ms=25, step_ms=10), dict(num_channels=32, window_size_ms=25, window_step_ms=10)),
])
def test_other_op_attributes(kws_lib, kw_o, kw_g):
    from multilingual_kws_b200.frontend import MicroFrontend
    pcm = synthetic_pcm(9, cfg_id=3)
    got = gpu_u16(MicroFrontend(**kw_g), pcm)
    assert np.array_equal(got, FrontendOracle(**kw_o).features_u16(pcm))


@pytest.mark.parametrize("hop", [320, 1600])
def test_streaming_frame_reuse_equals_per_window_recompute(fe, orc, hop):
    """batch_streaming_analysis.py:108-115 recomputes the frontend per window from a zero state; the
    frame-reuse path must give identical windows."""
    T = 16000 * 6 + 123
    x = synthetic_stream(T, cfg_id=5)
    st = fe.stream_prepare(torch.from_numpy(x).cuda())
    W = fe.stream_num_windows(T, 16000, hop)
    assert W == len(range(0, T - 16000, hop))
    a = st.windows(16000, hop, 0, W // 2)
    b = st.windows(16000, hop, W // 2, W - W // 2)
    got = torch.cat([a, b]).cpu().numpy()
    wins = np.stack([x[o:o + 16000] for o in range(0, T - 16000, hop)])
    assert np.array_equal(got, orc.features(wins, threads=8))


def test_size_independent_properties_full_batch(fe):
    """At bench size (B=8192): determinism, clip independence (permutation), zero clip -> zero rows."""
    pcm = synthetic_pcm(512, cfg_id=9)
    big = np.tile(pcm, (16, 1))
    big[100] = 0
    d = torch.from_numpy(big).cuda()
    a = fe.forward(d, raw_u16=True).cpu().numpy().astype(np.int32)
    perm = np.random.default_rng(0).permutation(big.shape[0])
    b = fe.forward(d[torch.from_numpy(perm).cuda()], raw_u16=True).cpu().numpy().astype(np.int32)
    assert np.array_equal(a[perm], b)
    assert a[100].max() == 0
    assert np.array_equal(a[:512][101:], a[512:1024][101:])      # tiled clips give tiled features
    assert a.sum() == a[perm].sum()
