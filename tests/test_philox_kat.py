"""Known-answer test of the counter-based generator behind the on-the-fly Theta: Philox4x32-10
against the vectors published with Random123 (kat_vectors, `philox4x32 10 ...`), on the host
(CPU test: a host-only entry point of the library, no GPU needed) and on the device."""
import numpy as np
import pytest

# counter (4 words), key (2 words) -> output (4 words)
KAT = [
    ([0x00000000] * 4, [0x00000000] * 2, [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def _pack():
    inp = np.array([c + k for c, k, _ in KAT], dtype=np.uint32)
    exp = np.array([o for _, _, o in KAT], dtype=np.uint32)
    return inp, exp


def test_philox_host_known_answers():
    from rla4mor_b200._lib import check, lib
    inp, exp = _pack()
    out = np.zeros_like(exp)
    check(lib().rla_philox4x32_10_host(inp.ctypes.data, len(inp), out.ctypes.data), "rla_philox4x32_10_host")
    assert np.array_equal(out, exp)


@pytest.mark.gpu
def test_philox_device_known_answers_and_host_device_agreement():
    import torch
    from rla4mor_b200._lib import check, lib, stream_ptr
    inp, exp = _pack()
    rs = np.random.RandomState(0)
    rnd = rs.randint(0, 2 ** 32, size=(4096, 6), dtype=np.uint64).astype(np.uint32)
    allin = np.concatenate([inp, rnd])
    host = np.zeros((len(allin), 4), dtype=np.uint32)
    check(lib().rla_philox4x32_10_host(allin.ctypes.data, len(allin), host.ctypes.data), "host")
    d_in = torch.from_numpy(allin.view(np.int32)).cuda()
    d_out = torch.empty((len(allin), 4), dtype=torch.int32, device="cuda")
    check(lib().rla_philox4x32_10_device(d_in.data_ptr(), len(allin), d_out.data_ptr(), stream_ptr()), "device")
    dev = d_out.cpu().numpy().view(np.uint32)
    assert np.array_equal(dev[:3], exp)
    assert np.array_equal(dev, host)


@pytest.mark.gpu
def test_theta_words_are_the_philox_stream():
    """Rademacher Theta (kind 1) is a pure function of the Philox words: bit j of the block
    (seed, row, col // 128) -- checked against the host generator, i.e. the same on any device."""
    import torch
    from rla4mor_b200 import dense
    from rla4mor_b200._lib import check, lib
    seed, rows, cols = 0x1234567890abcdef, 5, 512
    th = dense.theta_materialize(seed, dense.KIND_RADEMACHER, 1.0, rows, cols).cpu().numpy()
    for r in range(rows):
        for blk in range(cols // 128):
            inp = np.array([[blk, r, 0, 1, seed & 0xffffffff, seed >> 32]], dtype=np.uint32)
            out = np.zeros((1, 4), dtype=np.uint32)
            check(lib().rla_philox4x32_10_host(inp.ctypes.data, 1, out.ctypes.data), "host")
            bits = np.concatenate([(int(w) >> np.arange(32)) & 1 for w in out[0]])
            assert np.array_equal(th[r, blk * 128:(blk + 1) * 128], np.where(bits == 1, -1.0, 1.0))
