"""CPU tests of the host-side mirror of the reference interface: hdf5 hand-off layout, parameter
mapping, the shared optimizer step counter, copy-dropout mask, path helpers, sharding + gather."""
import os

import numpy as np
import pytest
import torch

from deeplabv3plus_augmented_superresolution_b200 import hdf5_lite, sharding, utils
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.optimizer import Optimizer
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.superresolution import Superresolution
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts import superres_utils as SU


def _write_case(path, n=6, mode="argmax", with_max=False):
    rng = np.random.RandomState(0)
    cm = [rng.rand(8, 8, 1).astype(np.float32) for _ in range(n)]
    f = hdf5_lite.File(path, "w")
    f.create_dataset("class_masks", data=cm)
    if with_max:
        f.create_dataset("max_masks", data=[c * 2 for c in cm])
    f.create_dataset("angles", data=rng.rand(n).astype(np.float32))
    f.create_dataset("shifts", data=rng.rand(n, 2).astype(np.float32))
    f.attrs["filename"] = "2007_000032"
    f.attrs["mode"] = mode
    f.attrs["angle_max"] = 0.15
    f.attrs["shift_max"] = 80
    f.close()
    return cm


def test_hdf5_layout_roundtrip(tmp_path):
    p = str(tmp_path / "2007_000032.hdf5")
    cm = _write_case(p, with_max=True, mode="slice_max")
    raw = open(p, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0                 # superblock v0, as h5py's default
    assert int.from_bytes(raw[40:48], "little") == len(raw)               # end-of-file address
    assert raw.count(b"TREE") == 1 and raw.count(b"SNOD") == 1 and raw.count(b"HEAP") == 1 and raw.count(b"GCOL") == 1
    f = hdf5_lite.File(p, "r")
    assert sorted(f) == ["angles", "class_masks", "max_masks", "shifts"]
    assert f["class_masks"].shape == (6, 8, 8, 1) and f["class_masks"].dtype == np.float32
    assert f["shifts"].shape == (6, 2) and f["angles"].shape == (6,)
    np.testing.assert_array_equal(f["class_masks"][:4], np.stack(cm)[:4])
    assert f.attrs["filename"] == "2007_000032" and isinstance(f.attrs["mode"], str) and f.attrs["mode"] == "slice_max"
    assert f.attrs["angle_max"] == 0.15 and f.attrs["shift_max"] == 80 and isinstance(f.attrs["shift_max"], int)
    assert SU.check_hdf5_validity(f, num_aug=6) and not SU.check_hdf5_validity(f, num_aug=7)
    with pytest.raises(KeyError):
        f["nope"]
    f.close()
    with pytest.raises(hdf5_lite.Hdf5FormatError):
        bad = tmp_path / "bad.hdf5"
        bad.write_bytes(b"not hdf5 at all" * 10)
        hdf5_lite.File(str(bad), "r")


def test_path_helpers(tmp_path):
    for name in ("2008_000003", "2007_000129", "2007_000032"):
        (tmp_path / f"{name}.hdf5").write_bytes(b"")
    (tmp_path / "notes.txt").write_text("x")
    got = [os.path.basename(p) for p in SU.list_precomputed_data_paths(str(tmp_path), sort=True)]
    assert got == ["2007_000032.hdf5", "2007_000129.hdf5", "2008_000003.hdf5"]
    lst = tmp_path / "list.txt"
    lst.write_text("2007_000129\n2007_000032\n")
    assert [os.path.basename(p) for p in SU.get_img_paths(str(lst), "/imgs")] == ["2007_000032.jpg", "2007_000129.jpg"]
    assert SU.normalize_coefficients({"a": 1.0, "b": 3.0}) == {"a": 0.25, "b": 0.75}
    x = np.array([[1.0, 3.0]], np.float32)
    np.testing.assert_array_equal(SU.min_max_normalization(x, 0.0, 1.0), [[0.0, 1.0]])
    np.testing.assert_array_equal(SU.min_max_normalization(np.full((2, 2), 4.0, np.float32), 0.0, 1.0), np.zeros((2, 2)))


def test_optimizer_and_params_mapping():
    o = Optimizer(optimizer="adam", learning_rate=1e-3, amsgrad=True, lr_scheduler=True, decay_steps=60, decay_rate=0.3)
    assert o.kind == "adam" and o.iterations == 0
    assert Optimizer(optimizer="something-else").kind == "adam"        # reference falls through to Adam (optimizer.py:36-41)
    assert abs(float(o.lr_decay(60)) - 3e-4) < 1e-9
    s = Superresolution(lambda_df=1.0, lambda_tv=0.3, lambda_L2=0.7, lambda_L1=0.0, num_iter=300, num_aug=100, optimizer=o,
                        feature_size=(128, 128))
    p = s._solve_params(step_offset=600)
    c = p.to_c()
    assert (c.num_iter, c.optimizer, c.amsgrad, c.lr_scheduler, c.step_offset) == (300, 0, 1, 1, 600)
    assert abs(c.lambda_tv - 0.3) < 1e-7 and abs(c.decay_rate - 0.3) < 1e-7 and c.decay_steps == 60.0
    with pytest.raises(Exception, match="must provide an instance of the Optimizer"):
        Superresolution(1, 1, 1, 0)._check_optimizer()
    Superresolution(1, 1, 1, 0, optimizer=o, output_size=(512, 512))._check_sizes(64, 64)       # the (64,64) default: x8 is supported
    with pytest.raises(NotImplementedError):
        Superresolution(1, 1, 1, 0, optimizer=o, output_size=(512, 512))._check_sizes(100, 100)   # non-integer ratio


def test_copy_dropout_mask_is_frozen_like_a_traced_function():
    o = Optimizer()
    s = Superresolution(1, 1, 1, 0, num_aug=10, optimizer=o, copy_dropout=0.35)
    np.random.seed(5)
    m1 = s._dropout_keep(10)
    m2 = s._dropout_keep(10)           # second call must not redraw (mask is baked in at trace time, :47-53)
    assert m1.sum() == 7 and np.array_equal(m1, m2)
    np.random.seed(5)
    ref = np.full(10, True); ref[:3] = False; np.random.shuffle(ref)
    np.testing.assert_array_equal(m1.astype(bool), ref)
    assert Superresolution(1, 1, 1, 0, num_aug=10, optimizer=o)._dropout_keep(10) is None


def test_resize_and_iou_helpers():
    a = np.arange(16, dtype=np.float32).reshape(4, 4, 1)
    up = utils._resize(a, (8, 8), "bilinear")
    ref = torch.nn.functional.interpolate(torch.from_numpy(a[None, :, :, 0])[None], size=(8, 8), mode="bilinear", align_corners=False)[0, 0].numpy()
    np.testing.assert_allclose(up[..., 0], ref, atol=1e-6)
    np.testing.assert_array_equal(utils._resize(a, (8, 8), "nearest")[::2, ::2], a)
    t = np.array([[8, 8, 0, 3]]); p = np.array([[8, 0, 0, 0]])
    assert utils.compute_IoU(t, p, img_size=(1, 4), class_id=8) == 0.5
    assert utils.compute_IoU(t, p, img_size=(1, 4), class_id=8, include_bg=True) == pytest.approx((0.5 + 2 / 3) / 2)
    assert np.isnan(utils.compute_IoU(np.zeros((1, 4)), np.zeros((1, 4)), img_size=(1, 4), class_id=8))
    assert utils.create_mask(np.array([[[0.1, 0.9, 0.9]]])).tolist() == [[[1]]]


def test_shard_bounds_cover_everything_once():
    for n, w in ((500, 8), (500, 1), (7, 8), (64, 3), (0, 4)):
        seen = []
        for r in range(w):
            a, b = sharding.shard_bounds(n, w, r)
            seen += list(range(a, b))
        assert seen == list(range(n))
    assert sharding.shard_bounds(500, 8, 7) == (441, 500)


def _gather_worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    a, b = sharding.shard_bounds(n_total, world, rank)
    local = torch.stack([torch.full((4, 4), i, dtype=torch.uint8) for i in range(a, b)]) if b > a else torch.zeros((0, 4, 4), dtype=torch.uint8)
    out = sharding.gather_masks(local, n_total)
    if rank == 0:
        q.put(out.numpy())
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5, 8])
def test_gather_masks_world_size_2_gloo(n_total):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    [p.start() for p in procs]
    out = q.get(timeout=120)
    [p.join(timeout=120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert out.shape == (n_total, 4, 4)
    np.testing.assert_array_equal(out[:, 0, 0], np.arange(n_total))


# ---------------------------------------------------------------------------------------------------------------
# hdf5: a file assembled by hand, byte by byte from the HDF5 File Format Specification, the way libhdf5 (behind h5py's
# default libver) lays one out -- NOT with hdf5_lite's writer -- with structures that writer never emits: a non-zero
# superblock base address (user block), NIL padding messages, object-modification-time and fill-value messages, an
# object-header continuation block, a version-3 attribute message, a version-2 dataspace, a compact dataset, a
# two-level group B-tree and a global heap with several objects.  h5py itself is absent (no wheel, no index), so a
# file written by real h5py cannot be committed; this pins the reader to the published format instead of to its sibling.
# ---------------------------------------------------------------------------------------------------------------
def _assemble_spec_file(arrays, attrs_str, attr_f64, attr_i64, userblock=512):
    import struct
    U = 0xFFFFFFFFFFFFFFFF
    pad8 = lambda b: b + b"\0" * (-len(b) % 8)
    msg = lambda t, body, fl=0: struct.pack("<HHB3x", t, len(pad8(body)), fl) + pad8(body)
    f32 = struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    f64 = struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    i64 = struct.pack("<BBBBI", 0x10, 0x08, 0, 0, 8) + struct.pack("<HH", 0, 64)
    vstr = struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + struct.pack("<BBBBI", 0x10, 0, 0, 0, 1) + struct.pack("<HH", 0, 8)
    space_v1 = lambda shape: struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", d) for d in shape)
    space_v2 = lambda shape: struct.pack("<BBBB", 2, len(shape), 0, 1 if shape else 0) + b"".join(struct.pack("<Q", d) for d in shape)

    blob = bytearray(b"\0" * 2048)                      # superblock + root structures are patched in at the end
    def put(b, align=8):
        while len(blob) % align:
            blob.append(0)
        a = len(blob)
        blob.extend(b)
        return a

    # global heap: object 1..n = the strings, then the free-space object 0
    gobjs, gbody = {}, b""
    for i, (k, v) in enumerate(attrs_str.items(), start=1):
        raw = v.encode() + b"\0"
        gobjs[k] = (i, len(raw))
        gbody += struct.pack("<HHIQ", i, 1, 0, len(raw)) + pad8(raw)
    gsize = 4096
    gbody += struct.pack("<HHIQ", 0, 0, 0, gsize - 16 - len(gbody) - 16)
    gcol = put(b"GCOL" + struct.pack("<B3xQ", 1, gsize) + gbody + b"\0" * (gsize - 16 - len(gbody)))

    # datasets: raw data first, then their object headers
    names = sorted(arrays)
    heads = {}
    for n in names:
        a = np.ascontiguousarray(arrays[n], np.float32)
        fill = msg(0x0005, struct.pack("<BBBB", 2, 2, 0, 0))                                     # fill value v2, undefined
        mtime = msg(0x0012, struct.pack("<B3xI", 1, 1639000000))                                 # object modification time
        if n == "angles":                                                                        # a compact dataset, v2 dataspace
            layout = msg(0x0008, struct.pack("<BBH", 3, 0, a.nbytes) + a.tobytes())
            msgs = [msg(0x0001, space_v2(a.shape)), msg(0x0003, f32, 1), fill, layout, mtime, msg(0x0000, b"\0" * 8)]
            heads[n] = put(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, sum(map(len, msgs))) + b"".join(msgs))
        else:
            data = put(a.tobytes())
            layout = msg(0x0008, struct.pack("<BBQQ", 3, 1, data - userblock, a.nbytes))
            # first block holds dataspace + datatype + a continuation; the rest lives in a continuation block elsewhere
            cont_msgs = [fill, layout, mtime, msg(0x0000, b"\0" * 16)]
            cont = put(b"".join(cont_msgs))
            first = [msg(0x0001, space_v1(a.shape)), msg(0x0003, f32, 1),
                     msg(0x0010, struct.pack("<QQ", cont - userblock, sum(map(len, cont_msgs))))]
            heads[n] = put(struct.pack("<BBHII4x", 1, 0, len(first) + len(cont_msgs), 1, sum(map(len, first))) + b"".join(first))

    # local heap with the link names, two SNODs under a level-1 TREE
    heap_data = bytearray(b"\0" * 8)
    offs = {}
    for n in names:
        offs[n] = len(heap_data)
        heap_data += pad8(n.encode() + b"\0")
    heap_data += b"\0" * 64
    hd = put(bytes(heap_data))
    heap = put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), len(heap_data) - 64, hd - userblock))
    def snod(group):
        ents = b"".join(struct.pack("<QQII16x", offs[n], heads[n] - userblock, 0, 0) for n in group)
        return put(b"SNOD" + struct.pack("<BBH", 1, 0, len(group)) + ents + b"\0" * (40 * (8 - len(group))))
    half = (len(names) + 1) // 2
    s1, s2 = snod(names[:half]), snod(names[half:])
    leaf = lambda s, lo, hi: put(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, U, U) + struct.pack("<QQQ", lo, s - userblock, hi) + b"\0" * 512)
    t1, t2 = leaf(s1, 0, offs[names[half - 1]]), leaf(s2, offs[names[half - 1]], offs[names[-1]])
    top = put(b"TREE" + struct.pack("<BBHQQ", 0, 1, 2, U, U) +
              struct.pack("<QQQQQ", 0, t1 - userblock, offs[names[half - 1]], t2 - userblock, offs[names[-1]]) + b"\0" * 512)

    # root object header: symbol table message + the four attributes (v1, v3 with an encoding byte, v1, v1)
    def attr_v1(name, dtype, space, data):
        nm = name.encode() + b"\0"
        return msg(0x000C, struct.pack("<BxHHH", 1, len(nm), len(dtype), len(space)) + pad8(nm) + pad8(dtype) + pad8(space) + data)
    def attr_v3(name, dtype, space, data):
        nm = name.encode() + b"\0"
        return msg(0x000C, struct.pack("<BBHHHB", 3, 0, len(nm), len(dtype), len(space), 1) + nm + dtype + space + data)
    vref = lambda k: struct.pack("<IQI", gobjs[k][1], gcol - userblock, gobjs[k][0])
    keys = list(attrs_str)
    rmsgs = [msg(0x0011, struct.pack("<QQ", top - userblock, heap - userblock)),
             attr_v1(keys[0], vstr, space_v1(()), vref(keys[0])),
             msg(0x0000, b"\0" * 24),
             attr_v3(keys[1], vstr, space_v2(()), vref(keys[1])),
             attr_v1(attr_f64[0], f64, space_v1(()), struct.pack("<d", attr_f64[1])),
             attr_v1(attr_i64[0], i64, space_v1(()), struct.pack("<q", attr_i64[1]))]
    root = put(struct.pack("<BBHII4x", 1, 0, len(rmsgs), 1, sum(map(len, rmsgs))) + b"".join(rmsgs))

    # superblock v0 at the start of the user-block-shifted address space
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", userblock, U, len(blob) - userblock, U)
    sb += struct.pack("<QQII", 0, root - userblock, 1, 0) + struct.pack("<QQ", top - userblock, heap - userblock)
    blob[userblock:userblock + len(sb)] = sb
    # libhdf5 looks for the signature at 0, 512, 1024, ...; hdf5_lite reads the base address from the superblock at offset 0,
    # so the user block here is expressed as the superblock's base-address field with the superblock itself at offset 0
    out = bytearray(blob)
    out[0:len(sb)] = sb
    return bytes(out)


def test_hdf5_reader_on_a_spec_level_file_it_did_not_write(tmp_path):
    rng = np.random.RandomState(11)
    arrays = {"class_masks": rng.rand(5, 6, 7, 1).astype(np.float32), "max_masks": rng.rand(5, 6, 7, 1).astype(np.float32),
              "angles": rng.rand(5).astype(np.float32), "shifts": rng.rand(5, 2).astype(np.float32)}
    raw = _assemble_spec_file(arrays, {"filename": "2008_000123", "mode": "slice_max"}, ("angle_max", 0.15), ("shift_max", 80))
    p = tmp_path / "2008_000123.hdf5"
    p.write_bytes(raw)
    f = hdf5_lite.File(str(p), "r")
    assert sorted(f) == sorted(arrays)
    for k, v in arrays.items():
        assert f[k].shape == v.shape and f[k].dtype == np.float32
        np.testing.assert_array_equal(f[k][:4], v[:4])
    assert f.attrs["filename"] == "2008_000123" and f.attrs["mode"] == "slice_max"
    assert f.attrs["angle_max"] == 0.15 and f.attrs["shift_max"] == 80
    assert SU.check_hdf5_validity(f, num_aug=5) and not SU.check_hdf5_validity(f, num_aug=6)
    f.close()


_H5PY_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "h5py_fixture.hdf5")


@pytest.mark.skipif(not os.path.exists(_H5PY_FIXTURE), reason="no h5py-written fixture in this repo yet (oracle/tf_crosscheck.py --write-hdf5; "
                    "h5py cannot be installed here: profiles/r02_tf_install_attempt.log)")
def test_hdf5_reader_on_h5py_fixture():
    """A file written by real h5py with the reference's own calls (augmentation_utils.py:123-136), when one is committed."""
    f = hdf5_lite.File(_H5PY_FIXTURE, "r")
    want = np.load(_H5PY_FIXTURE + ".class_masks.npy")
    np.testing.assert_array_equal(f["class_masks"][:len(want)], want)
    assert f.attrs["filename"] == "2007_000032" and f.attrs["mode"] == "argmax" and f.attrs["angle_max"] == 0.15 and f.attrs["shift_max"] == 80
    f.close()
