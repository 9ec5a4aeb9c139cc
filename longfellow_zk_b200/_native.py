"""ctypes loader for liblongfellow_b200.so.  There is no Python or CPU fallback:
if the CUDA library is missing or no device is present, calls raise."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.environ.get("LF_LIB_PATH") or os.path.join(HERE, "liblongfellow_b200.so")  # LF_LIB_PATH: tuning builds

_lib = None


class LongfellowError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"longfellow_b200 error {code}: {msg}")
        self.code = code


class CircuitInfo(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "ninputs", "npub_in", "nl", "nterms", "kbytes", "witness_bytes", "rng_bytes",
        "max_proof_bytes", "block_enc", "block", "dblock", "block_ext", "nrow", "r", "w", "nwrow",
        "nqtriples", "nreq", "nw", "sumcheck_alg_bytes", "sumcheck_mults", "total_mults", "sha_compressions",
        "rs_mults", "eval_mults", "ligero_mults", "merkle_compressions", "lfc1_bytes",
        "rng_sample_bytes", "rng_redraw_bytes", "rng_redraw_cap", "sumcheck_proof_elts",
        "flat_rounds", "flat_eval_alg_bytes", "flat_eval_mults", "flat_bind_alg_bytes", "flat_bind_mults")]


class Transcript(C.Structure):
    """lf_transcript: the transcript state that crosses the C ABI"""
    _fields_ = [("h", C.c_uint32 * 8), ("buf", C.c_uint32 * 16), ("len", C.c_uint64), ("nblock", C.c_uint64),
                ("rdptr", C.c_uint32), ("have_prf", C.c_uint32), ("saved", C.c_uint32 * 4)]


EXPORTS = [
    "lf_ctx_create", "lf_ctx_destroy", "lf_ctx_synchronize", "lf_last_error", "lf_version",
    "lf_elt_mul", "lf_rs_interpolate", "lf_rs_interpolate_dev", "lf_merkle_commit",
    "lf_circuit_upload", "lf_circuit_free", "lf_circuit_get_info", "lf_zk_prove_batch",
    "lf_zk_prove_batch_dev", "lf_zk_debug_fetch", "lf_ctx_launch_count", "lf_microbench",
    "lf_circuit_set_profiling", "lf_circuit_get_stage_ms", "lf_fft", "lf_fft_time",
    "lf_zk_commit_batch", "lf_zk_prove_committed_batch", "lf_transcript_init", "lf_transcript_write_bytes",
    "lf_transcript_challenge_bytes", "lf_zk_rng_consumed", "lf_circuit_verify_id", "lf_zk_verify_batch", "lf_zk_verify_committed_batch", "lf_zk_verify_set_fault", "lf_fft_time_rows", "lf_rs_time", "lf_circuit_get_kernel_ms",
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise ImportError(
                f"{LIB} is missing: build it with `python -m longfellow_zk_b200.build` "
                "(the CUDA extension is mandatory; there is no CPU fallback)")
        L = C.CDLL(LIB)
        L.lf_last_error.restype = C.c_char_p
        L.lf_version.restype = C.c_char_p
        L.lf_ctx_launch_count.restype = C.c_uint64
        L.lf_ctx_launch_count.argtypes = [C.c_void_p]
        L.lf_ctx_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.lf_ctx_destroy.argtypes = [C.c_void_p]
        L.lf_ctx_synchronize.argtypes = [C.c_void_p]
        L.lf_elt_mul.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.lf_rs_interpolate.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p,
                                        C.c_size_t]
        L.lf_rs_interpolate_dev.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p,
                                            C.c_size_t, C.c_size_t]
        L.lf_merkle_commit.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.lf_circuit_upload.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, C.c_size_t,
                                        C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p)]
        L.lf_circuit_free.argtypes = [C.c_void_p]
        L.lf_circuit_get_info.argtypes = [C.c_void_p, C.POINTER(CircuitInfo)]
        L.lf_zk_prove_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                        C.c_void_p]
        L.lf_zk_prove_batch_dev.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                            C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                            C.c_void_p]
        L.lf_zk_debug_fetch.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t,
                                        C.POINTER(C.c_size_t)]
        L.lf_circuit_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.lf_circuit_get_stage_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_size_t]
        L.lf_fft.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int]
        L.lf_fft_time.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
        L.lf_fft_time_rows.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
        L.lf_rs_time.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                                 C.POINTER(C.c_double)]
        L.lf_circuit_get_kernel_ms.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_size_t)]
        L.lf_microbench.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.lf_zk_commit_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.POINTER(Transcript), C.c_void_p, C.c_void_p]
        L.lf_zk_prove_committed_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(Transcript),
                                                  C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.lf_transcript_init.argtypes = [C.POINTER(Transcript), C.c_char_p, C.c_size_t]
        L.lf_transcript_init.restype = None
        L.lf_transcript_write_bytes.argtypes = [C.POINTER(Transcript), C.c_char_p, C.c_size_t]
        L.lf_transcript_write_bytes.restype = None
        L.lf_transcript_challenge_bytes.argtypes = [C.POINTER(Transcript), C.c_void_p, C.c_size_t]
        L.lf_transcript_challenge_bytes.restype = None
        L.lf_circuit_verify_id.argtypes = [C.c_void_p, C.c_void_p]
        L.lf_zk_rng_consumed.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.lf_zk_verify_set_fault.argtypes = [C.c_void_p, C.c_int]
        L.lf_zk_verify_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                         C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.lf_zk_verify_committed_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                                   C.c_void_p, C.POINTER(Transcript), C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise LongfellowError(rc, lib().lf_last_error().decode())
