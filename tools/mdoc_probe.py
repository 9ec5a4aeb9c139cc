"""Upload the two circuits of the mdoc circuit file (zstd, lib/circuits/mdoc/circuits/8d0792...,
ZkSpec v7, 1 attribute: block_enc 4151 / 4096) and report sizes and upload times."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import longfellow_zk_b200 as lf
z = C.CDLL("libzstd.so.1")
z.ZSTD_decompress.restype = C.c_size_t
z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
data = open(os.path.join(ROOT, "tests/golden/mdoc/circuit_v7_1attr.zst"), "rb").read()
cap = z.ZSTD_getFrameContentSize(data, len(data))
buf = C.create_string_buffer(cap)
n = z.ZSTD_decompress(buf, cap, data, len(data))
raw = buf.raw[:n]
ctx = lf.Context(0)
t = time.time(); sig = lf.Circuit(ctx, 1, raw, block_enc=4096); print("sig upload %.1f s" % (time.time() - t), sig.info, flush=True)
off = sig.info["lfc1_bytes"]
t = time.time(); h = lf.Circuit(ctx, 4, raw[off:], block_enc=4151); print("hash upload %.1f s" % (time.time() - t), h.info, flush=True)
import torch
print("device memory in use (GB): %.1f" % ((torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9))
