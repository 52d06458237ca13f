"""TensorFlow checkpoint container + import/export (tf_checkpoint.py) — CPU only.  Known answers come from the public format
definitions (CRC-32C check value, leveldb table magic, snappy's format description); everything else is writer -> reader."""
import os
import struct

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def ck(tb):
    return tb.tf_checkpoint


def test_crc32c_known_answers_native_and_python(tb, ck):
    assert ck.crc32c(b"123456789") == 0xE3069283                        # the CRC-32C check value (RFC 3720 B.4)
    assert ck.crc32c(b"\x00" * 32) == 0x8A9136AA and ck.crc32c(b"\xff" * 32) == 0x62A8AB43    # RFC 3720 test patterns
    rng = np.random.RandomState(0)
    blob = rng.bytes(70001)                                             # >= 4096 bytes goes through libtaco_b200
    lib = tb.capi.load()
    native = int(lib.taco_crc32c(blob, len(blob), 0))
    c = 0xFFFFFFFF
    for b in blob:
        c = int(ck._CRC_TABLE[(c ^ b) & 0xFF]) ^ (c >> 8)
    assert native == (c ^ 0xFFFFFFFF) == ck.crc32c(blob)
    part = int(lib.taco_crc32c(blob[:12345], 12345, 0))                 # incremental form
    assert int(lib.taco_crc32c(blob[12345:], len(blob) - 12345, part)) == native
    assert ck._mask(0) == 0xa282ead8


def test_bundle_round_trip_multi_block(tmp_path, ck):
    rng = np.random.RandomState(1)
    tensors = {"scope_%03d/w/kernel" % i: rng.randn(3, i % 5 + 1).astype(np.float32) for i in range(150)}
    tensors.update({"global_step": np.array(1234, dtype=np.int32), "a/scalar": np.array(0.5, dtype=np.float32),
                    "a/i64": np.arange(7, dtype=np.int64), "a/empty": np.zeros((0, 4), dtype=np.float32),
                    "a/f64": rng.randn(2, 2, 2), "a/flag": np.array([True, False])})
    prefix = str(tmp_path / "model.ckpt-1234")
    ck.write_bundle(prefix, tensors, data_crc=True)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xdb4775248b80fb57       # leveldb/TF table magic
    got = ck.read_bundle(prefix, verify=True, verify_data=True)
    assert sorted(got) == sorted(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    # a flipped bit in a data block is caught by the block checksum; one in the tensor data by the entry checksum
    bad = bytearray(raw); bad[10] ^= 0x40
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        ck.read_bundle(prefix)
    open(prefix + ".index", "wb").write(raw)
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read()); data[100] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="data checksum"):
        ck.read_bundle(prefix, verify_data=True)
    with pytest.raises(ValueError, match="bad table magic"):
        open(prefix + ".index", "wb").write(raw[:-1] + b"\x00")
        ck.read_bundle(prefix)


def test_snappy_block_decoding(ck):
    # literal "abcd", copy(offset 4, len 8) with a 1-byte offset tag, literal "XY", copy with a 2-byte offset (len 5, offset 6)
    comp = bytes([19]) + bytes([(4 - 1) << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4]) + bytes([(2 - 1) << 2]) + b"XY" + \
        bytes([((5 - 1) << 2) | 2, 6, 0])
    assert ck._snappy_decompress(comp) == b"abcdabcdabcdXY" + b"abcdX"
    long_lit = bytes(range(256)) * 2
    comp = ck._put_varint(len(long_lit)) + bytes([61 << 2]) + struct.pack("<H", len(long_lit) - 1) + long_lit
    assert ck._snappy_decompress(comp) == long_lit
    with pytest.raises(ValueError):
        ck._snappy_decompress(bytes([4, ((4 - 4) << 2) | 1, 9]))


@pytest.mark.parametrize("over,S", [(dict(), 1), (dict(model_type="deepvoice"), 3), (dict(model_type="simple", attention_type="bah_norm"), 2)])
def test_export_import_state_round_trip(tmp_path, tb, ck, over, S):
    hp = tb.hparams.override(reduction_factor=5, **over)
    specs = tb.params.param_specs(hp, S)
    layout = tb.params.make_layout(specs)
    named = tb.params.init_params(hp, S, seed=3, randomize_bn_state=True)
    params, bn = tb.params.flatten(named, layout)
    g = torch.Generator().manual_seed(1)
    mask = torch.zeros_like(params)
    for s in specs:                                                       # padding between tensors is not part of the state
        if s.trainable:
            mask[layout.offsets[s.name]:layout.offsets[s.name] + s.numel] = 1
    state = dict(params=params, bn_state=bn, adam_m=torch.randn(params.shape, generator=g) * mask,
                 adam_v=torch.rand(params.shape, generator=g) * mask, global_step=0, adam_step=4321,
                 names=[s.name for s in specs], hparams=hp.values(), num_speakers=S)
    prefix = str(tmp_path / "run" / "model.ckpt-0")
    ck.export_state(prefix, state, hp, S)
    bundle = ck.read_bundle(prefix, verify_data=True)
    assert bundle["global_step"].dtype == np.int32 and int(bundle["global_step"]) == 0
    assert "model/inference/embedding" in bundle and "model/inference/embedding/Adam_1" in bundle
    assert "model/inference/encoder_cbhg/conv_bank/conv1d_16/batch_normalization/moving_variance" in bundle
    assert "model/inference/encoder_cbhg/conv_bank/conv1d_16/batch_normalization/moving_variance/Adam" not in bundle
    if hp.attention_type == "bah_mon":
        key = [k for k in bundle if k.endswith("attention_score_bias")][0]
        assert bundle[key].shape == ()                                    # a TF scalar, a 1-vector here
    back = ck.import_state(prefix + ".index", hp, S)                      # any of the file names resolves to the prefix
    for k in ("params", "bn_state", "adam_m", "adam_v"):
        assert torch.equal(back[k], state[k]), k
    assert back["global_step"] == 0 and back["adam_step"] == 4321 and back["names"] == state["names"]
    assert ck.is_tf_checkpoint(prefix) and ck.load_any(prefix, hp, S)["adam_step"] == 4321
    # the directory scan of the reference (models/__init__.py:10-17) finds it
    assert tb.get_most_recent_checkpoint(os.path.dirname(prefix)) == prefix
    # wrong hyper-parameters: a clear error naming what is missing
    with pytest.raises(KeyError, match="lacks"):
        ck.import_state(prefix, tb.hparams.override(reduction_factor=5, model_type="deepvoice"), 5 if S != 5 else 6) if hp.model_type != "deepvoice" \
            else ck.import_state(prefix, tb.hparams.override(reduction_factor=5, attention_type="bah_norm", model_type="deepvoice"), S)
    with pytest.raises(ValueError, match="expects"):
        ck.import_state(prefix, tb.hparams.override(reduction_factor=4, **over), S)
