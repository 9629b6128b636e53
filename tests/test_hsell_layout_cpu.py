"""CPU check of the host arithmetic of the hsell streams (pygrank_b200.graph.stream_layout): piece
numbering, chunk descriptors and block boundaries against a round-by-round brute-force model."""
import numpy as np
import pytest
import torch

from pygrank_b200.graph import stream_layout


def brute(rounds, base, CH):
    B, S = rounds.shape
    g0 = np.zeros(B * S, dtype=np.int64)
    p0 = np.zeros(B * S, dtype=np.int64)
    pieces = np.zeros(B * S, dtype=np.int64)
    chunk_first, chunk_mask, chunk_begin = [], [], []
    g = 0            # global round
    piece = base     # next piece number
    for b in range(B):
        assert g % CH == 0
        chunk_begin.append(g // CH)
        for s in range(S):
            r = int(rounds[b, s])
            u = b * S + s
            g0[u] = g
            p0[u] = piece          # only meaningful when r > 0
            for k in range(r):
                if g % CH == 0:
                    chunk_first.append(piece)
                    chunk_mask.append(0)
                last_of_unit = k == r - 1
                if last_of_unit:
                    chunk_mask[-1] |= 1 << (g % CH)
                if last_of_unit or g % CH == CH - 1:
                    piece += 1
                    pieces[u] += 1
                g += 1
        while g % CH:              # padding rounds of the block: they close one junk piece at the chunk end
            if g % CH == CH - 1:
                piece += 1
            g += 1
    chunk_begin.append(g // CH)
    return g0, p0, pieces, np.array(chunk_first, dtype=np.int64), np.array(chunk_mask, dtype=np.int64), g // CH, \
        piece - base, np.array(chunk_begin, dtype=np.int64)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("CH", [4, 32])
def test_stream_layout_matches_brute_force(seed, CH):
    rng = np.random.default_rng(seed)
    B, S = int(rng.integers(1, 5)), int(rng.integers(1, 40))
    rounds = rng.integers(0, 3 * CH, size=(B, S))
    rounds[rng.random((B, S)) < 0.4] = 0
    if seed == 0:
        rounds[:] = 0                       # nothing at all
    if seed == 1:
        rounds[0, :] = CH                   # every unit ends exactly on a chunk boundary
    base = int(rng.integers(0, 1000))
    g0, p0, pieces, desc, n_chunks, n_parts, chunk_begin = stream_layout(torch.from_numpy(rounds), base, CH)
    bg0, bp0, bpieces, bfirst, bmask, bn_chunks, bn_parts, bbegin = brute(rounds, base, CH)
    ex = rounds.reshape(-1) > 0
    assert n_chunks == bn_chunks and n_parts == bn_parts
    assert np.array_equal(g0.numpy()[ex], bg0[ex])
    assert np.array_equal(p0.numpy()[ex], bp0[ex])
    assert np.array_equal(pieces.numpy(), bpieces)
    assert np.array_equal(chunk_begin.numpy(), bbegin)
    d = desc.numpy().astype(np.int64)
    if n_chunks:
        assert np.array_equal(d[:, 0], bfirst)
        assert np.array_equal(d[:, 1] & (2 ** CH - 1), bmask & (2 ** CH - 1))


def test_hsell_shape_and_config_defaults(monkeypatch):
    """block size / block count rules (no GPU needed: pgb_hsell_max_block_cols is pure arithmetic)."""
    from pygrank_b200 import graph
    for k in ("PGB_HSELL_BLOCK_COLS", "PGB_HSELL_BLOCKS"):
        monkeypatch.delenv(k, raising=False)
    H, K = graph.hsell_shape(torch.float32, 1, 1 << 24)
    assert H == 32768 and K == 80
    H, K = graph.hsell_shape(torch.float64, 1, 1 << 24)
    assert H == 16384 and K == 160
    H, K = graph.hsell_shape(torch.float32, 8, 1 << 24)         # 8 ranks x 16.8 M rows: 134 M columns
    assert H == 32768 and K == 256
    monkeypatch.setenv("PGB_HSELL_BLOCKS_CAP", "128")
    assert graph.hsell_shape(torch.float32, 8, 1 << 24)[1] == 128
    monkeypatch.delenv("PGB_HSELL_BLOCKS_CAP")
    H, K = graph.hsell_shape(torch.float32, 1, 1000)            # tiny graph: one partial block
    assert K == 1
    H, K = graph.hsell_shape(torch.float32, 2, 5000)            # multi-segment blocks must be full
    assert K == 0
    H, K = graph.hsell_shape(None, 1, 1 << 24, elem_bytes=graph.PANEL_ELEM_BYTES)   # panel form: 16 bytes per node
    assert H == 8192 and K == 320
    monkeypatch.setenv("PGB_HSELL_BLOCK_COLS", "1000000")        # clamped to what shared memory holds
    assert graph.hsell_shape(None, 1, 1 << 24, elem_bytes=16)[0] == ((232448 - 64) // 16 - 1) & ~63
    H, _ = graph.hsell_shape(torch.float32, 1, 1 << 24)
    assert 0 < H <= 65535 and H % 4 == 0
