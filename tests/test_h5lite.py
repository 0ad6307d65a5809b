"""CPU: h5lite - the subset of HDF5 the reference's input / output files use (vamp_1.0/vpspectrum.py:58-64, 528-538).

The reader is checked against files written by h5py itself: the reference's fixtures under /root/reference (build
container only; their arrays are also in tests/golden/simba.npz, extracted independently by byte offset) and a
32-byte-exact fragment of one of them kept here.  The writer is checked through the reader and against the message
bytes real h5py wrote."""
import os

import numpy as np
import pytest

from vamp_b200 import h5lite

REF = "/root/reference/vamp_1.0/data"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# the four messages h5py / libhdf5 wrote for a contiguous float64[1000] dataset (simba_H1215.h5, object header at 0x320):
# dataspace, datatype, fill value, layout (its address field zeroed here)
H5PY_F64_1000 = bytes.fromhex(
    "0100180000000000" "0101010000000000" "e803000000000000" "e803000000000000"
    "0300180001000000" "11203f0008000000" "0000400034" "0b0034ff03000000000000"
    "0500080001000000" "0202020100000000"
    "0800180000000000" "0301" "0000000000000000" "401f000000000000" "000000000000")


def test_round_trip_all_supported_types(tmp_path):
    rng = np.random.default_rng(0)
    data = {"wavelength": rng.uniform(1200, 1300, 1000), "flux": rng.uniform(size=1000).astype(np.float32),
            "region_pixels": np.array([[3, 40], [60, 99]], dtype=np.int64), "difficult_fit": np.array(True),
            "chi_squared": np.array(1.25), "empty": np.array([]), "region_0_flux": rng.uniform(size=(3, 37)),
            "counts": np.arange(7, dtype=np.uint16), "i32": np.array([-5, 6], dtype=np.int32)}
    for i in range(40):                                   # more entries than a default symbol node holds
        data["region_%d_wave" % i] = rng.uniform(size=5 + i)
    path = str(tmp_path / "t.h5")
    h5lite.write_h5(path, data)
    back = h5lite.read_h5(path)
    assert sorted(back) == sorted(data)
    for k, v in data.items():
        want = np.asarray(v)
        if want.dtype == np.bool_:
            want = want.astype(np.uint8)
        assert back[k].dtype == want.dtype and back[k].shape == want.shape, k
        np.testing.assert_array_equal(back[k], want)
    assert os.path.getsize(path) == h5lite._Reader(open(path, "rb").read()).eof


def test_writer_emits_the_messages_h5py_wrote(tmp_path):
    """Same dataset shape and type as the fixture's: the object header messages are the bytes libhdf5 produced."""
    path = str(tmp_path / "one.h5")
    h5lite.write_h5(path, {"flux": np.zeros(1000)})
    buf = open(path, "rb").read()
    r = h5lite._Reader(buf)
    (name, hdr), = r.group_entries(*r.u("QQ", 80))
    assert name == "flux"
    msgs = r.messages(hdr)
    assert [m[0] for m in msgs] == [0x0001, 0x0003, 0x0005, 0x0008]
    start, end = msgs[0][1] - 8, msgs[3][1] + msgs[3][2]
    got = bytearray(buf[start:end])
    addr_at = msgs[3][1] + 2 - start
    got[addr_at: addr_at + 8] = b"\0" * 8                 # the data address differs from file to file
    assert bytes(got) == H5PY_F64_1000
    # superblock, root entry and root header sit where h5py puts them (offsets 0, 56, 96; B-tree at 136)
    assert buf[:8] == h5lite.SIGNATURE and r.root_header == 96 and r.u("QQ", 80)[0] == 136


def test_rejects_what_it_does_not_understand(tmp_path):
    p = tmp_path / "bad.h5"
    p.write_bytes(b"not hdf5 at all" * 10)
    with pytest.raises(h5lite.H5Unsupported):
        h5lite.read_h5(str(p))
    with pytest.raises(h5lite.H5Unsupported):
        h5lite.write_h5(str(tmp_path / "x.h5"), {"s": np.array(["a", "b"])})
    with pytest.raises(h5lite.H5Unsupported):
        h5lite.write_h5(str(tmp_path / "x.h5"), {"a/b": np.zeros(3)})
    good = tmp_path / "v2.h5"
    h5lite.write_h5(str(good), {"a": np.zeros(3)})
    b = bytearray(good.read_bytes())
    b[8] = 2                                              # a newer superblock
    good.write_bytes(bytes(b))
    with pytest.raises(h5lite.H5Unsupported):
        h5lite.read_h5(str(good))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference fixtures live in the build container only")
def test_reads_the_reference_fixtures_written_by_h5py():
    g = np.load(os.path.join(GOLDEN, "simba.npz"))
    for key, nreg in (("simba_CII1036", 4), ("simba_H1215", 3)):
        d = h5lite.read_h5(os.path.join(REF, key + ".h5"))
        assert sorted(d) == ["density_col", "flux", "noise", "tau", "temp", "velocity", "wavelength"]
        assert all(v.shape == (1000,) and v.dtype == np.float64 for v in d.values())
        assert np.all(np.diff(d["wavelength"]) > 0) and 0.0 < d["noise"].min()
        # the golden regions were cut from arrays read by raw byte offset (oracle/load_reference.py): same numbers
        bounds = {"simba_CII1036": [[677, 714], [720, 765], [774, 791], [932, 955]],
                  "simba_H1215": [[672, 716], [711, 762], [767, 796]]}[key]
        for j, (s, e) in enumerate(bounds):
            np.testing.assert_array_equal(np.flip(d["flux"][s:e], 0), g["%s/r%d/flux" % (key, j)])
            np.testing.assert_array_equal(np.flip(d["noise"][s:e], 0), g["%s/r%d/noise" % (key, j)])
    other = h5lite.read_h5(os.path.join(REF, "ramses_ray.h5"))           # a third h5py-written file, other contents
    assert len(other) == 17 and all(k.startswith("grid/") and v.shape == (178,) for k, v in other.items())   # a nested group


def test_vpspectrum_reads_and_writes_reference_format(tmp_path):
    """``VPspectrum(line, spectrum_file=...)`` and ``write_file`` without h5py (vpspectrum.py:58-64, 528-538)."""
    from vamp_b200.vpspectrum import VPspectrum
    rng = np.random.default_rng(2)
    wave = 1215.0 + 0.02 * np.arange(300)
    flux, noise = rng.uniform(0.2, 1.0, 300), np.full(300, 0.02)
    src = str(tmp_path / "sight.h5")
    h5lite.write_h5(src, {"wavelength": wave, "flux": flux, "noise": noise, "tau": -np.log(flux)})
    sp = VPspectrum(1215.0, spectrum_file=src, out_folder=str(tmp_path) + "/")
    np.testing.assert_array_equal(sp.flux_array, flux)
    np.testing.assert_array_equal(sp.wavelength_array, wave)
    sp.params = {"b": np.array([20.0, 31.0]), "N": np.array([1e13, 3e13]), "EW": np.array([0.1, 0.2])}
    sp.flux_model = {"total": np.ones(300), "chi_squared": np.array([1.1]), "region_pixels": [[10, 60]],
                     "difficult_fit": False, "region_0_flux": np.ones((2, 50))}
    sp.output_filename = str(tmp_path / "sight_gauss_")
    sp.write_file()
    p = h5lite.read_h5(sp.output_filename + "params.h5")
    m = h5lite.read_h5(sp.output_filename + "flux_model.h5")
    np.testing.assert_array_equal(p["N"], sp.params["N"])
    assert m["region_pixels"].tolist() == [[10, 60]] and m["region_0_flux"].shape == (2, 50) and m["difficult_fit"] == 0
