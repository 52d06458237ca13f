"""TensorFlow ``model.ckpt-<step>`` checkpoints (tensor-bundle V2) without TensorFlow: read, write, import, export.

The reference saves and restores with ``tf.train.Saver`` (train.py:175,189-205,242-244; synthesizer.py:56-58), which in
TF 1.x writes the V2 "tensor bundle": ``<prefix>.index`` (an SSTable: key = variable name, value = BundleEntryProto) and
``<prefix>.data-00000-of-00001`` (raw little-endian tensor bytes).  A user switching over brings such checkpoints along,
so this module restates the two formats (TF sources: core/util/tensor_bundle/tensor_bundle.cc, core/lib/io/{table,
format,block}.cc, core/protobuf/tensor_bundle.proto — all public, none of it needs TF at run time):

  * ``read_bundle(prefix)`` / ``write_bundle(prefix, tensors)`` — the container;
  * ``import_state(prefix, hp, num_speakers)`` — variables, Adam slots (``<var>/Adam``, ``<var>/Adam_1``), ``beta1_power``
    and ``global_step`` -> the state dict ``models.Tacotron.load_state_dict`` takes;
  * ``export_state(prefix, state, hp, num_speakers)`` — the way back.

Variable names come from ``tf_names.tf_to_ours``.  Validation status: the container code is tested against itself (writer
-> reader, CRCs, multi-block indexes, snappy blocks) and against the format constants of the TF sources; no checkpoint
written by a real TensorFlow was available in this environment.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Optional, Tuple

import numpy as np

from . import params as P
from .tf_names import tf_to_ours

_MAGIC = 0xdb4775248b80fb57          # table/format.h kTableMagicNumber
_DT = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8"), 10: np.dtype("bool"),
       4: np.dtype("u1"), 6: np.dtype("i1"), 5: np.dtype("<i2"), 17: np.dtype("<u2"), 19: np.dtype("<f2")}   # types.proto DataType
_DT_INV = {v: k for k, v in _DT.items()}


# ---- crc32c (Castagnoli), masked as leveldb does ---------------------------------------------------------------------
def _make_crc_table():
    tbl = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
        tbl[i] = c
    return tbl


_CRC_TABLE = _make_crc_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    """Native (libtaco_b200: taco_crc32c, slicing-by-8) when the library is there, else a byte-wise Python loop."""
    if len(data) >= 4096:
        try:
            from . import capi
            return int(capi.load().taco_crc32c(bytes(data), len(data), crc))
        except Exception:
            pass
    c = crc ^ 0xFFFFFFFF
    tbl = _CRC_TABLE
    for b in data:
        c = int(tbl[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _mask(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF


# ---- varints / protobuf wire format -------------------------------------------------------------------------------
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if b < 0x80:
            return out, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _parse_proto(buf: bytes) -> Dict[int, list]:
    """Field number -> list of raw values (ints for varint / fixed, bytes for length-delimited)."""
    out: Dict[int, list] = {}
    pos = 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]; pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _field(tag: int, wt: int, payload: bytes) -> bytes:
    return _put_varint((tag << 3) | wt) + payload


def _entry_proto(dtype: int, shape, shard: int, offset: int, size: int, crc: int) -> bytes:
    """BundleEntryProto {1 dtype, 2 shape {2 dim {1 size}}, 3 shard_id, 4 offset, 5 size, 6 crc32c (fixed32)}."""
    dims = b"".join(_field(2, 2, (lambda d: _put_varint(len(d)) + d)(_field(1, 0, _put_varint(int(s))))) for s in shape)
    out = _field(1, 0, _put_varint(dtype)) + _field(2, 2, _put_varint(len(dims)) + dims)
    if shard:
        out += _field(3, 0, _put_varint(shard))
    if offset:
        out += _field(4, 0, _put_varint(offset))
    out += _field(5, 0, _put_varint(size)) + _field(6, 5, struct.pack("<I", crc))
    return out


# ---- snappy (block decompression only; TF's bundle writer stores uncompressed blocks, other writers may not) ----------
def _snappy_decompress(buf: bytes) -> bytes:
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]; pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little"); pos += nb
            ln += 1
            out += buf[pos:pos + ln]; pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]; pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8); pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little"); pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy block")
        for _ in range(ln):                       # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch")
    return bytes(out)


# ---- SSTable ----------------------------------------------------------------------------------------------------
def _read_block(f: bytes, offset: int, size: int, verify: bool) -> bytes:
    body, trailer = f[offset:offset + size], f[offset + size:offset + size + 5]
    if verify:
        want = struct.unpack("<I", trailer[1:5])[0]
        if _mask(crc32c(body + trailer[:1])) != want:
            raise ValueError("block checksum mismatch at offset %d" % offset)
    if trailer[0] == 0:
        return body
    if trailer[0] == 1:
        return _snappy_decompress(body)
    raise ValueError("unknown block compression %d" % trailer[0])


def _block_entries(block: bytes):
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        unshared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + unshared]; pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _read_table(path: str, verify: bool = True) -> Dict[bytes, bytes]:
    with open(path, "rb") as fh:
        f = fh.read()
    if len(f) < 48 or struct.unpack("<Q", f[-8:])[0] != _MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad table magic)" % path)
    footer = f[-48:]
    _mo, p = _get_varint(footer, 0)
    _ms, p = _get_varint(footer, p)
    io, p = _get_varint(footer, p)
    isz, p = _get_varint(footer, p)
    out: Dict[bytes, bytes] = {}
    for _k, handle in _block_entries(_read_block(f, io, isz, verify)):
        bo, q = _get_varint(handle, 0)
        bs, q = _get_varint(handle, q)
        for k, v in _block_entries(_read_block(f, bo, bs, verify)):
            out[k] = v
    return out


def _build_block(items, restart_interval: int = 16) -> bytes:
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _write_table(path: str, items, block_entries: int = 64) -> None:
    items = sorted(items)
    f = bytearray()

    def emit(block: bytes) -> bytes:
        off = len(f)
        f.extend(block + b"\x00" + struct.pack("<I", _mask(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    index = []
    for i in range(0, len(items), block_entries):
        chunk = items[i:i + block_entries]
        index.append((chunk[-1][0], emit(_build_block(chunk))))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index, restart_interval=1))
    footer = meta + idx
    f.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC))
    with open(path, "wb") as fh:
        fh.write(bytes(f))


# ---- tensor bundle ------------------------------------------------------------------------------------------------
def read_bundle(prefix: str, verify: bool = True, verify_data: bool = False) -> Dict[str, np.ndarray]:
    """{variable name: array} of a V2 checkpoint ``<prefix>.index`` + ``<prefix>.data-?????-of-?????``."""
    table = _read_table(prefix + ".index", verify)
    header = _parse_proto(table.get(b"", b""))
    num_shards = header.get(1, [1])[0]
    if header.get(2, [0])[0] != 0:
        raise ValueError("big-endian checkpoints are not supported")
    shards = {}
    out: Dict[str, np.ndarray] = {}
    for key, val in table.items():
        if key == b"":
            continue
        e = _parse_proto(val)
        if 7 in e:
            raise ValueError("partitioned (sliced) variable %s is not supported" % key.decode())
        dt = e.get(1, [0])[0]
        if dt not in _DT:
            raise ValueError("variable %s has unsupported dtype enum %d" % (key.decode(), dt))
        shape = []
        for sp in e.get(2, []):
            for dim in _parse_proto(sp).get(2, []):
                shape.append(_signed64(_parse_proto(dim).get(1, [0])[0]))
        shard, offset, size = e.get(3, [0])[0], e.get(4, [0])[0], e.get(5, [0])[0]
        if shard not in shards:
            shards[shard] = np.memmap("%s.data-%05d-of-%05d" % (prefix, shard, num_shards), dtype=np.uint8, mode="r")
        raw = shards[shard][offset:offset + size]
        if verify_data and 6 in e and _mask(crc32c(raw.tobytes())) != e[6][0]:
            raise ValueError("data checksum mismatch for %s" % key.decode())
        arr = np.frombuffer(raw.tobytes(), dtype=_DT[dt])
        if int(np.prod(shape, dtype=np.int64)) != arr.size:
            raise ValueError("variable %s: %d bytes do not fill shape %s" % (key.decode(), size, shape))
        out[key.decode()] = arr.reshape(shape)
    return out


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], data_crc: bool = False) -> None:
    """Writes ``<prefix>.index`` and ``<prefix>.data-00000-of-00001`` (one shard, no compression, as tf.train.Saver does).
    data_crc=True fills the per-tensor crc32c field (slow in pure Python; TF verifies it on restore, so turn it on for files
    that go back to TensorFlow)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = [(b"", _field(1, 0, _put_varint(1)) + _field(3, 2, (lambda v: _put_varint(len(v)) + v)(_field(1, 0, _put_varint(1)))))]
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        for name in sorted(tensors):
            a = np.asarray(tensors[name])                      # (ascontiguousarray would turn a scalar into a 1-vector)
            dt = _DT_INV.get(a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype)
            if dt is None:
                raise ValueError("dtype %s of %s cannot be stored" % (a.dtype, name))
            raw = a.astype(_DT[dt], copy=False).tobytes(order="C")
            fh.write(raw)
            crc = _mask(crc32c(raw)) if data_crc else 0
            items.append((name.encode(), _entry_proto(dt, a.shape, 0, offset, len(raw), crc)))
            offset += len(raw)
    _write_table(prefix + ".index", items)


# ---- reference model <-> this repo ----------------------------------------------------------------------------------
def _resolve_prefix(path: str) -> str:
    for suf in (".index", ".meta"):
        if path.endswith(suf):
            return path[:-len(suf)]
    if ".data-" in os.path.basename(path):
        return path[:path.rindex(".data-")]
    return path


def is_tf_checkpoint(path: str) -> bool:
    return os.path.exists(_resolve_prefix(path) + ".index")


def import_state(path: str, hp, num_speakers: int = 1, scope: str = "model") -> dict:
    """A reference checkpoint -> the state dict of ``models.Tacotron.load_state_dict``.  Raises KeyError listing every
    variable of the model the checkpoint lacks (wrong hyper-parameters / speaker count), ValueError on a shape mismatch."""
    import torch
    prefix = _resolve_prefix(path)
    bundle = read_bundle(prefix)
    table = tf_to_ours(hp, num_speakers, prefix=scope + "/inference/")
    specs = P.param_specs(hp, num_speakers)
    layout = P.make_layout(specs)
    by_name = {s.name: s for s in specs}
    missing = [k for k in table if k not in bundle]
    if missing:
        raise KeyError("checkpoint %s lacks %d variable(s) of this model (hyper-parameters / num_speakers mismatch?): %s"
                       % (prefix, len(missing), ", ".join(sorted(missing)[:8])))
    named, m, v = {}, {}, {}
    for tf_name, ours in table.items():
        spec = by_name[ours]
        arr = bundle[tf_name]
        if int(arr.size) != spec.numel or (arr.ndim > 0 and tuple(arr.shape) != tuple(spec.shape) and arr.size != 1):
            raise ValueError("%s has shape %s in the checkpoint, the model expects %s" % (tf_name, arr.shape, spec.shape))
        named[ours] = torch.from_numpy(np.array(arr, dtype=np.float32).reshape(spec.shape))
        if spec.trainable:
            for slot, dst in (("/Adam", m), ("/Adam_1", v)):
                s = bundle.get(tf_name + slot)
                dst[ours] = torch.zeros(spec.shape) if s is None else torch.from_numpy(np.array(s, dtype=np.float32).reshape(spec.shape))
    params, bn_state = P.flatten(named, layout)
    zeros = {s.name: torch.zeros(s.shape) for s in specs if not s.trainable}
    adam_m, _ = P.flatten({**zeros, **m}, layout)
    adam_v, _ = P.flatten({**zeros, **v}, layout)
    global_step = int(bundle["global_step"]) if "global_step" in bundle else 0
    adam_step = global_step
    # beta<i>_power = beta<i> ** (updates + 1), float32: beta1_power underflows to 0 after ~10^3 updates, beta2_power after
    # ~10^5 — by then the bias correction is 1 to within rounding and global_step is as good
    for suffix, beta in (("beta2_power", hp.adam_beta2), ("beta1_power", hp.adam_beta1)):
        keys = [k for k in bundle if k.endswith(suffix)]
        if keys and 1e-30 < float(bundle[keys[0]]) < 1.0:
            adam_step = max(0, int(round(np.log(float(bundle[keys[0]])) / np.log(beta))) - 1)
            break
    return dict(params=params, bn_state=bn_state, adam_m=adam_m, adam_v=adam_v, global_step=global_step, adam_step=adam_step,
                names=[s.name for s in specs], hparams=hp.values(), num_speakers=num_speakers, source=prefix)


def export_state(prefix: str, state: dict, hp, num_speakers: int = 1, scope: str = "model", data_crc: bool = True) -> None:
    """The way back: a state dict -> ``<prefix>.index`` / ``.data-00000-of-00001`` a ``tf.train.Saver`` of the reference's
    graph restores (variables, Adam slots, beta powers, global_step)."""
    specs = P.param_specs(hp, num_speakers)
    layout = P.make_layout(specs)
    if [s.name for s in specs] != list(state["names"]):
        raise RuntimeError("state dict tensor inventory does not match these hyper-parameters")
    views = P.views(state["params"].cpu(), state["bn_state"].cpu(), layout)
    mv = P.views(state["adam_m"].cpu(), state["bn_state"].cpu(), layout)
    vv = P.views(state["adam_v"].cpu(), state["bn_state"].cpu(), layout)
    out: Dict[str, np.ndarray] = {}
    for tf_name, ours in tf_to_ours(hp, num_speakers, prefix=scope + "/inference/").items():
        spec = layout.spec(ours)
        scalar = tf_name.endswith(("attention_score_bias", "attention_g"))
        shp = () if scalar else spec.shape
        out[tf_name] = views[ours].numpy().reshape(shp)
        if spec.trainable:
            out[tf_name + "/Adam"] = mv[ours].numpy().reshape(shp)
            out[tf_name + "/Adam_1"] = vv[ours].numpy().reshape(shp)
    t = int(state.get("adam_step", state["global_step"]))
    out[scope + "/optimizer/beta1_power"] = np.array(hp.adam_beta1 ** (t + 1), dtype=np.float32)
    out[scope + "/optimizer/beta2_power"] = np.array(hp.adam_beta2 ** (t + 1), dtype=np.float32)
    out["global_step"] = np.array(int(state["global_step"]), dtype=np.int32)
    write_bundle(prefix, out, data_crc=data_crc)


def load_any(path: str, hp=None, num_speakers: Optional[int] = None) -> dict:
    """``model.ckpt-<step>.pt`` (this repo) or a TensorFlow checkpoint prefix / .index file (the reference)."""
    if path.endswith(".pt") and os.path.exists(path):
        import torch
        return torch.load(path, map_location="cpu", weights_only=False)
    if is_tf_checkpoint(path):
        if hp is None or num_speakers is None:
            raise ValueError("importing a TensorFlow checkpoint needs the hyper-parameters and the speaker count")
        return import_state(path, hp, num_speakers)
    raise FileNotFoundError(path)
