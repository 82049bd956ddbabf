"""Reader for Torch7's binary serialization (torch.save / torch.DiskFile():binary()) -- enough of it to
bring a model trained with the reference into this library without Torch7: the network file written by
torch/lib/save_model.lua and the `_mconf.bin` table beside it (torch/lib/load_model.lua).

Format (torch7/File.lua, the un-vendored Torch7 distro the reference runs on): every value starts with an
int32 type tag -- 0 nil, 1 number (float64), 2 string (int32 length + bytes), 3 table, 4 torch object,
5 boolean, 6/7/8 functions.  Tables and torch objects carry an int32 index first; an index seen before is
a back-reference.  A torch object continues with a version string ("V 1"), its class name and the class's
own payload: tensors write ndim, sizes, strides (int64 each), storage offset (1-based) and their storage
object; storages write their length and the raw elements; every other class (nn modules, nngraph nodes)
writes one table with its fields.  torch.Cuda* tensors / storages are written like their Float versions.
Host-side only, no GPU involved.
"""
import struct

import numpy as np

_STORAGE_DTYPES = {
    "torch.FloatStorage": np.float32, "torch.CudaStorage": np.float32, "torch.DoubleStorage": np.float64,
    "torch.CudaDoubleStorage": np.float64, "torch.LongStorage": np.int64, "torch.CudaLongStorage": np.int64,
    "torch.IntStorage": np.int32, "torch.CudaIntStorage": np.int32, "torch.ByteStorage": np.uint8,
    "torch.CudaByteStorage": np.uint8, "torch.CharStorage": np.int8, "torch.ShortStorage": np.int16,
    "torch.HalfStorage": np.float16, "torch.CudaHalfStorage": np.float16,
}


class T7Object:
    """A torch class instance that is not a tensor / storage: `cls` and its field table `fields`."""

    def __init__(self, cls, fields):
        self.cls = cls
        self.fields = fields

    def __getitem__(self, key):
        return self.fields[key]

    def get(self, key, default=None):
        return self.fields.get(key, default) if isinstance(self.fields, dict) else default

    def __repr__(self):
        return "T7Object(%s)" % self.cls


class _ObjKey:
    """A table / object used as a table key (nngraph keeps node -> node maps): hashable by identity."""

    def __init__(self, obj):
        self.obj = obj

    def __hash__(self):
        return id(self.obj)

    def __eq__(self, other):
        return isinstance(other, _ObjKey) and other.obj is self.obj


class _Reader:
    def __init__(self, data):
        self.d = data
        self.o = 0
        self.memo = {}

    def _unpack(self, fmt, size):
        v = struct.unpack_from(fmt, self.d, self.o)
        self.o += size
        return v[0]

    def int32(self):
        return self._unpack("<i", 4)

    def int64(self):
        return self._unpack("<q", 8)

    def float64(self):
        return self._unpack("<d", 8)

    def string(self):
        n = self.int32()
        s = self.d[self.o:self.o + n].decode("latin-1")
        self.o += n
        return s

    def value(self):
        tag = self.int32()
        if tag == 0:
            return None
        if tag == 1:
            v = self.float64()
            return int(v) if (v == v and abs(v) < 2 ** 53 and v == int(v)) else v
        if tag == 2:
            return self.string()
        if tag == 5:
            return self.int32() == 1
        if tag == 3:
            idx = self.int32()
            if idx in self.memo:
                return self.memo[idx]
            t = {}
            self.memo[idx] = t
            n = self.int32()
            for _ in range(n):
                k = self.value()
                if isinstance(k, (dict, T7Object, np.ndarray)):
                    k = _ObjKey(k)
                t[k] = self.value()
            return t
        if tag == 4:
            idx = self.int32()
            if idx in self.memo:
                return self.memo[idx]
            version = self.string()
            cls = self.string() if version.startswith("V ") else version
            return self._torch_object(idx, cls)
        if tag in (6, 7, 8):                     # functions: dumped bytecode (+ upvalues); skipped
            if tag in (7, 8):
                idx = self.int32()
                if idx in self.memo:
                    return self.memo[idx]
                self.memo[idx] = "<function>"
            self.string()                        # the dump
            if tag in (7, 8):
                self.value()                     # upvalues table
            return "<function>"
        raise ValueError("torch7 file: unknown type tag %d at offset %d" % (tag, self.o - 4))

    def _torch_object(self, idx, cls):
        if cls in _STORAGE_DTYPES:
            n = self.int64()
            dt = np.dtype(_STORAGE_DTYPES[cls])
            a = np.frombuffer(self.d, dtype=dt, count=n, offset=self.o).copy()
            self.o += n * dt.itemsize
            self.memo[idx] = a
            return a
        if cls.endswith("Tensor") and cls.startswith("torch."):
            nd = self.int32()
            size = [self.int64() for _ in range(nd)]
            stride = [self.int64() for _ in range(nd)]
            offset = self.int64() - 1
            storage = self.value()
            if storage is None or nd == 0:
                a = np.zeros(size if nd else (0,), np.float32)
            else:
                a = np.lib.stride_tricks.as_strided(storage[offset:], shape=size,
                                                    strides=[s * storage.itemsize for s in stride]).copy()
            self.memo[idx] = a
            return a
        obj = T7Object(cls, None)
        self.memo[idx] = obj
        obj.fields = self.value()
        return obj


def load(path):
    """Deserialize the first value of a Torch7 binary file."""
    import sys
    with open(path, "rb") as f:
        data = f.read()
    limit = sys.getrecursionlimit()
    sys.setrecursionlimit(max(limit, 20000))         # nngraph models nest deeply (node -> children -> node ...)
    try:
        return _Reader(data).value()
    finally:
        sys.setrecursionlimit(limit)


def _walk(obj, seen, out):
    """Depth-first over tables / objects in insertion order, collecting convolution modules."""
    if id(obj) in seen:
        return
    if isinstance(obj, T7Object):
        seen.add(id(obj))
        if obj.cls.endswith("Convolution") and isinstance(obj.fields, dict) and "weight" in obj.fields:
            out.append(obj)
            return
        _walk(obj.fields, seen, out)
    elif isinstance(obj, dict):
        seen.add(id(obj))
        keys = list(obj.keys())
        ints = sorted(k for k in keys if isinstance(k, int))
        for k in ints + [k for k in keys if not isinstance(k, int)]:
            if isinstance(k, _ObjKey):
                _walk(k.obj, seen, out)
            _walk(obj[k], seen, out)


def conv_layers(model):
    """[(weight [cout][cin][kz][ky][kx], bias [cout]), ...] of a deserialized reference model, in forward
    order.  The reference's graphs are nngraph gModules: `forwardnodes` is already topologically sorted
    (nngraph/gmodule.lua), so the convolutions are taken from there; a plain nn.Sequential is walked in
    module order."""
    convs = []
    nodes = model.get("forwardnodes") if isinstance(model, T7Object) else None
    if nodes:
        for i in sorted(nodes):
            data = nodes[i].get("data") if isinstance(nodes[i], T7Object) else None
            mod = data.get("module") if isinstance(data, dict) else None
            if mod is not None:
                _walk(mod, set(), convs)
    else:
        _walk(model, set(), convs)
    layers = []
    for c in convs:
        w = np.asarray(c["weight"], np.float32)
        b = np.asarray(c["bias"], np.float32).reshape(-1)
        cout, cin = int(c["nOutputPlane"]), int(c["nInputPlane"])
        if "kT" in c.fields:                                     # Volumetric: kT x kH x kW
            k = (int(c["kT"]), int(c["kH"]), int(c["kW"]))
        else:                                                    # Spatial: kH x kW
            k = (1, int(c["kH"]), int(c["kW"]))
        layers.append((np.ascontiguousarray(w.reshape(cout, cin, *k)), np.ascontiguousarray(b)))
    return layers


def load_reference_model(model_path, mconf_path=None):
    """The pieces fluidnet_b200.model.ProjectionModel needs from a model saved by the reference
    (torch/lib/save_model.lua): {'is3D', 'layers', 'mconf'}.  `mconf_path` defaults to
    `<model_path>_mconf.bin` (torch/lib/load_model.lua)."""
    mconf = load(mconf_path or (model_path + "_mconf.bin"))
    model = load(model_path)
    if isinstance(model, dict) and "model" in model:
        model = model["model"]
    return {"is3D": bool(mconf.get("is3D")), "layers": conv_layers(model), "mconf": mconf}
