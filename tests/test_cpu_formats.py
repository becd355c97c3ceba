"""CPU tests: wire formats against the golden fixtures written by the reference's own Python
(tests/golden/make_golden.py), the Euler convention, the merge3d log contract."""
import os

import numpy as np
import pytest

from pyp_b200.formats import cistem, dump, mrc, statistics

G = os.path.join(os.path.dirname(__file__), "golden")


def test_cistem_read_matches_reference_readback():
    rows = cistem.read_parameters(os.path.join(G, "params_5x32.cistem"))
    want = np.load(os.path.join(G, "params_5x32_readback.npy"))  # what Parameters.from_file returned
    assert rows.dtype.itemsize == 128 and rows.size == 5
    for k, name in enumerate(cistem.ROW_DTYPE.names):  # column order cistem_star_file.py:596-628
        assert np.array_equal(rows[name].astype(np.float64), want[:, k]), name


def test_cistem_write_is_byte_identical(tmp_path):
    src = os.path.join(G, "params_5x32.cistem")
    rows = cistem.read_parameters(src)
    out = tmp_path / "x.cistem"
    cistem.write_parameters(str(out), rows)
    assert out.read_bytes() == open(src, "rb").read()
    assert len(out.read_bytes()) == 8 + 32 * 9 + 5 * 128


def test_cistem_extended_roundtrip(tmp_path):
    src = os.path.join(G, "params_5x32_extended.cistem")
    particles, tilts = cistem.read_extended(src)
    assert particles.dtype.itemsize == 48 and tilts.dtype.itemsize == 24  # Appendix B
    assert list(tilts["angle"]) == [-60.0, -57.0, -54.0]
    out = tmp_path / "x_extended.cistem"
    cistem.write_extended(str(out), particles, tilts)
    assert out.read_bytes() == open(src, "rb").read()
    assert cistem.extended_path("a/b_r01.cistem") == "a/b_r01_extended.cistem"


def test_cistem_merge_matches_reference():
    got = cistem.merge([os.path.join(G, "merge_a.cistem"), os.path.join(G, "merge_b.cistem")])
    want = cistem.read_parameters(os.path.join(G, "merge_ab.cistem"))
    assert got.tobytes() == want.tobytes()
    assert list(got["position_in_stack"]) == [1, 2, 3, 4, 5]


def test_cistem_rejects_broken_files(tmp_path):
    p = tmp_path / "bad.cistem"
    p.write_bytes(b"\x01\x00")
    with pytest.raises(ValueError):
        cistem.read_parameters(str(p))
    good = open(os.path.join(G, "params_5x32.cistem"), "rb").read()
    p.write_bytes(good[:-10])
    with pytest.raises(ValueError):
        cistem.read_parameters(str(p))
    bad = bytearray(good)
    bad[8:16] = (12345).to_bytes(8, "little")  # unknown column id
    p.write_bytes(bytes(bad))
    with pytest.raises(ValueError):
        cistem.read_parameters(str(p))


@pytest.mark.parametrize("name", ["volume_4x6x8", "stack_3x8x8"])
def test_mrc_read_and_header(tmp_path, name):
    h, data = mrc.read(os.path.join(G, name + ".mrc"))
    want = np.load(os.path.join(G, name + ".npy"))
    assert np.array_equal(np.asarray(data), want)
    assert (h["nz"], h["ny"], h["nx"]) == want.shape and h["mode"] == 2 and h["nsymbt"] == 0
    out = tmp_path / "w.mrc"
    mrc.write(str(out), want)
    a, b = np.frombuffer(out.read_bytes(), np.uint8), np.frombuffer(open(os.path.join(G, name + ".mrc"), "rb").read(), np.uint8)
    assert a.size == b.size
    diff = np.nonzero(a != b)[0]
    stats_bytes = set(range(76, 88)) | set(range(216, 220))  # amin/amax/amean/rms: float rounding only
    assert set(diff.tolist()) <= stats_bytes
    fa, fb = np.frombuffer(out.read_bytes()[:1024], "<f4"), np.frombuffer(b.tobytes()[:1024], "<f4")
    assert np.allclose(fa[[19, 20, 21, 54]], fb[[19, 20, 21, 54]], rtol=1e-5)


def test_mrc_slices_and_append(tmp_path):
    stack = np.load(os.path.join(G, "stack_3x8x8.npy"))
    _, d = mrc.read(os.path.join(G, "stack_3x8x8.mrc"), first=2, last=3)
    assert np.array_equal(np.asarray(d), stack[1:3])
    with pytest.raises(ValueError):
        mrc.read(os.path.join(G, "stack_3x8x8.mrc"), first=3, last=4)
    p = tmp_path / "s.mrc"
    mrc.write(str(p), stack[:1])
    mrc.append(str(p), stack[1:])
    h, d = mrc.read(str(p))
    assert h["nz"] == 3 and h["mz"] == 3 and np.array_equal(np.asarray(d), stack)
    with pytest.raises(ValueError):
        mrc.append(str(p), np.zeros((1, 4, 4), np.float32))


def test_euler_convention_matches_reference_decode(oracle):
    """geometry/core.py:222-247 decodes (psi, theta, phi) from pyp's left-handed matrix
    L(phi,theta,psi); the FREALIGN matrix the engine builds is M(psi,theta,phi) = L(-phi,-theta,-psi)
    (oracle/SEMANTICS.md §Euler).  Both directions are pinned on the golden vectors."""
    from pyp_b200 import synth

    g = np.load(os.path.join(G, "euler_decode.npy"))
    for row in g:
        L = row[:9].reshape(3, 3)
        psi, theta, phi = row[9:12]
        assert np.allclose(row[12:15], [psi, theta, phi], atol=1e-9)  # reference decode returns the inputs
        M = synth.euler_matrix(psi, theta, phi)
        # L is M with every angle negated
        assert np.allclose(L, synth.euler_matrix(-psi, -theta, -phi), atol=1e-12)
        assert np.allclose(oracle.euler_matrix(psi, theta, phi), M, atol=2e-6)


def test_merge3d_log_is_parsed_by_reference_slicing():
    st = np.zeros((9, 7))
    st[:, 0] = np.arange(9)
    st[1:, 1] = 16 * 1.35 / np.arange(1, 9)
    st[:, 2] = np.arange(9) / 16
    st[:, 3] = np.linspace(1, 0.1, 9)
    st[:, 4] = st[:, 3] * 0.9
    st[:, 5] = np.linspace(40, 1, 9)
    st[:, 6] = np.linspace(50, 1, 9)
    text = "banner line\n\n" + statistics.merge3d_log(st)
    table = statistics.parse_merge3d_log(text)  # frealign.py:2558-2567 arithmetic
    assert table.shape == (8, 7)
    assert np.allclose(table[:, 1], np.round(st[1:, 1], 2))  # column 1 = resolution
    assert np.allclose(table[:, 3], np.round(st[1:, 3], 4))  # column 3 = FSC


def test_statistics_file_roundtrip(tmp_path):
    st = np.zeros((5, 7))
    st[:, 0] = np.arange(5)
    st[1:, 1] = [100, 50, 33.3, 25]
    st[:, 5] = [0, 30, 10, 3, 1]
    p = tmp_path / "statistics_r01.txt"
    statistics.write_statistics(str(p), st)
    back = statistics.read_statistics(str(p))
    assert back.shape == (4, 7) and np.allclose(back[:, 1], st[1:, 1], atol=1e-4)
    w = statistics.ring_weights_from_statistics(back, 64, 1.35)
    assert w.shape == (65,) and np.all((w >= 0) & (w <= 1)) and w[1] > w[-1]


def test_dump_roundtrip(tmp_path):
    acc = np.random.default_rng(0).normal(size=(8, 8, 5, 4)).astype(np.float32)
    p = tmp_path / "x_map1_n1.mrc"
    dump.write(str(p), acc, 8, 1, 0, 1.35, 17)
    meta, back = dump.read(str(p))
    assert meta == {"box": 8, "pad": 1, "half": 0, "pixel_size": pytest.approx(1.35), "n_inserted": 17}
    assert np.array_equal(back, acc)
    assert dump.seed_paths("a/temp_map1_n.mrc", 3) == ["a/temp_map1_n1.mrc", "a/temp_map1_n2.mrc", "a/temp_map1_n3.mrc"]
    with pytest.raises(ValueError):
        dump.seed_paths("a/temp_map1.mrc", 3)
    p.write_bytes(b"x" * 100)
    with pytest.raises(ValueError):
        dump.read(str(p))


def test_star_tables_round_trip_and_reference_merge():
    """rows_{a,b}.star were written by formats/star.py; rows_star_merged_by_reference.npy is what the
    reference's merge_star (cistem_star_file.py:1398-1441) made of them."""
    from pyp_b200.formats import star

    rows = cistem.read_parameters(os.path.join(G, "params_5x32.cistem"))
    merged_ref = np.load(os.path.join(G, "rows_star_merged_by_reference.npy"))
    for k, name in enumerate(rows.dtype.names):
        assert np.array_equal(merged_ref[:, k].astype(rows.dtype[name]), rows[name]), name
    back = np.concatenate([star.read_star(os.path.join(G, "rows_a.star")), star.read_star(os.path.join(G, "rows_b.star"))])
    assert back.tobytes() == rows.tobytes()


def test_merge_with_film_id_matches_reference():
    want = np.load(os.path.join(G, "merge_filmid_by_reference.npy"))
    got = cistem.merge_with_film_id([os.path.join(G, "merge_a.cistem"), os.path.join(G, "merge_b.cistem")])
    assert got.size == want.shape[0] == 5 and list(got["image_is_active"]) == [0, 0, 0, 1, 1]
    for k, name in enumerate(got.dtype.names):
        assert np.array_equal(want[:, k].astype(got.dtype[name]), got[name]), name
    assert cistem.merge_with_film_id([]).size == 0
