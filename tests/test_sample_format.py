"""The training samples written by device self-play (tests/golden/selfplay_samples.bin, made on a B200 by
tests/golden/make_selfplay_fixture.py) against the reference's record format (src/training_data.rs:24-60,
python/train.py:16-20): decoded back into positions and checked through the oracle, and — where the reference
checkout is present — loaded with the reference's own dataset class."""
import os
import sys

import numpy as np
import pytest

from oracle import pyoracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "selfplay_samples.bin")
REF_PY = "/root/reference/python"


def records():
    rec = np.fromfile(FIXTURE, dtype=np.float32)
    assert rec.size % 5763 == 0
    return rec.reshape(-1, 5763)


def decode(planes):
    """Planes of the side-to-move view -> the position with WHITE to move that encodes to them (a black-to-move
    position and its colour-flipped twin have the same planes and the same STM-relative policy indices)."""
    b = O.Board()
    p = planes.reshape(17, 8, 8)
    for pl in range(12):
        bb = 0
        for row in range(8):
            for col in range(8):
                if p[pl, row, col] == 1.0:
                    bb |= 1 << ((7 - row) * 8 + col)      # white to move: row = 7 - rank (src/tensor.rs:29-37)
        b.pieces[pl] = bb
    for i in range(6):
        b.occ[0] |= b.pieces[i]
        b.occ[1] |= b.pieces[6 + i]
    b.w_to_move = 1
    ep = np.argwhere(p[12] == 1.0)
    b.en_passant = 0xFF if len(ep) == 0 else (7 - int(ep[0][0])) * 8 + int(ep[0][1])
    b.castling = sum(bit for bit, pl in ((1, 13), (2, 14), (4, 15), (8, 16)) if p[pl, 0, 0] == 1.0)
    for pl in (13, 14, 15, 16):
        assert (p[pl] == p[pl, 0, 0]).all()               # castling planes are broadcast
    return b


def test_records_decode_to_positions_whose_legal_moves_carry_the_policy():
    rec = records()
    assert rec.shape == (12, 5763)
    plies_with_captures = 0
    for r in rec:
        planes, material, flag, z, policy = r[:1088], r[1088], r[1089], r[1090], r[1091:]
        assert np.isin(planes, (0.0, 1.0)).all()
        b = decode(planes)
        assert bin(b.pieces[5]).count("1") == 1 and bin(b.pieces[11]).count("1") == 1     # one king each
        assert np.array_equal(O.board_to_planes(b), planes)                               # the encoder's own inverse
        legal = {O.move_to_index_stm(m, True) for m in O.legal_moves(b)}
        support = set(np.nonzero(policy)[0].tolist())
        assert support and support <= legal                     # visit shares only on legal moves of that position
        assert abs(float(policy.sum()) - 1.0) < 1e-5 and (policy >= 0).all()
        assert z in (-1.0, 0.0, 1.0) and flag == 1.0            # the exchange line always completes
        # dM of the sample = the oracle's exchange line (on the twin with white to move; the fixture holds no position
        # where mirroring the ranks reorders equal MVV-LVA captures)
        assert material == np.float32(O.forced_principal_exchange(b)[0])
        plies_with_captures += int(material != np.float32(O.pst_eval_cp(b) / 100.0))
    print("samples whose exchange line differs from the stand pat:", plies_with_captures)


@pytest.mark.skipif(not os.path.isdir(REF_PY), reason="reference checkout not present (GPU box)")
def test_the_reference_loader_reads_the_file(tmp_path):
    """python/train.py's ChessDataset, imported in place, on a directory holding our file."""
    import shutil
    shutil.copy(FIXTURE, tmp_path / "game_000.bin")
    sys.path.insert(0, REF_PY)
    try:
        import train as ref_train
    finally:
        sys.path.remove(REF_PY)
    assert ref_train.SAMPLE_SIZE_FLOATS == 5763
    ds = ref_train.ChessDataset(str(tmp_path), augment=False)
    rec = records()
    assert len(ds) == rec.shape[0]
    item = ds[3]
    tensors = [np.asarray(t) for t in item]
    assert tensors[0].shape == (17, 8, 8) and np.array_equal(tensors[0].reshape(-1), rec[3, :1088])
    flat = np.concatenate([np.asarray(t, dtype=np.float32).reshape(-1) for t in item])
    assert np.isclose(flat, rec[3, 1088]).any() and np.isclose(flat, rec[3, 1090]).any()     # material and value target
    pol = [t for t in tensors if t.size == 4672][0]
    assert np.array_equal(pol.reshape(-1), rec[3, 1091:])


@pytest.mark.skipif(not os.path.isdir(REF_PY), reason="reference checkout not present (GPU box)")
def test_the_reference_replay_buffer_samples_the_file(tmp_path):
    """python/replay_buffer.py (the training loop's data source, python/orchestrate.py), imported in place: add the
    game file, draw a batch, and find every drawn row among our records field for field."""
    import shutil
    games = tmp_path / "games"
    games.mkdir()
    shutil.copy(FIXTURE, games / "game_000.bin")
    sys.path.insert(0, REF_PY)
    try:
        import replay_buffer as ref_rb
    finally:
        sys.path.remove(REF_PY)
    rec = records()
    rb = ref_rb.ReplayBuffer(1000, str(tmp_path / "buffer"))
    assert rb.add_games(str(games)) == rec.shape[0] == rb.total_positions()
    np.random.seed(0)
    boards, scalars, values, policies = rb.sample_batch(32)
    assert boards.shape == (32, 17, 8, 8) and scalars.shape == (32, 2) and values.shape == (32, 1) and policies.shape == (32, 4672)
    for i in range(32):
        row = np.concatenate([boards[i].reshape(-1), scalars[i], values[i], policies[i]])
        assert (rec == row).all(axis=1).any()
