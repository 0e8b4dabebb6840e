"""Host-side mirror of the reference's counting interface on top of the C ABI.

Reference interface (bin/KPopCount.ml:20-64)::

    module KMerCounter (KIH: KMers.IntHash_t): sig
      val compute: ?verbose:bool -> linter:(string -> string) -> Files.ReadsIterate.t -> int -> string -> string -> unit
    end                                              (* store            max_results_size label  fname *)

``KMerCounter(k, content, max_results_size).compute(inputs, label, fname)`` keeps the argument meaning: ``inputs`` is
the ordered ``Files.ReadsIterate.t`` (``("fasta", path)``, ``("single-end", path)``, ``("paired-end", p1, p2)``), an empty
``label`` means one spectrum per sequence (-L), ``fname`` "" means "return the text" (the reference writes to stdout).
The functor's k range check (KMers.ml:264-267, 145-148) fires in the constructor, as it does at functor application.
Errors the reference dies of with an uncaught exception are raised as ``KPopCountError`` carrying the KPC_E_* code.
"""
import ctypes
import os

from . import _native as N


class KPopCountError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


class Content:
    """bin/KPopCount.ml:66-82"""
    DNA_ss, DNA_ds, Protein = N.KPC_DNA_SS, N.KPC_DNA_DS, N.KPC_PROTEIN
    _names = {"DNA-ss": DNA_ss, "DNA-single-stranded": DNA_ss, "DNA-ds": DNA_ds, "DNA-double-stranded": DNA_ds,
              "protein": Protein, "prot": Protein}

    @classmethod
    def of_string(cls, w):
        if w not in cls._names:
            raise ValueError(f"Invalid_content({w!r})")
        return cls._names[w]


def spectra_filename(prefix):
    """KMerDB.Spectra.make_filename (lib/KMerDB.ml:26-31)"""
    return prefix if prefix.startswith("/dev/") else prefix + ".KPopSpectra.txt"


_FORMATS = {"fasta": N.KPC_FASTA, "single-end": N.KPC_FASTQ_SE, "paired-end": N.KPC_FASTQ_PE}


class KMerCounter:
    def __init__(self, k=12, content=Content.DNA_ds, max_results_size=16777216, label="", device=0, lib=None, devices=None):
        """devices=[0, 1, ...]: several GPUs behind this one context (FASTQ inputs of dense-table runs are spread over
        them, the tables are added up on the first one at finish); default: the single GPU `device`."""
        self._lib = lib or N.load()
        self._ctx = ctypes.c_void_p()
        self.k, self.content, self.max_results_size, self.label = k, content, max_results_size, label
        ids = list(devices) if devices else [device]
        dev = (ctypes.c_int * len(ids))(*ids)
        rc = self._lib.kpc_create(ctypes.byref(self._ctx), k, content, max_results_size, label.encode(), len(ids), dev)
        if rc != N.KPC_OK:
            msg = self._lib.kpc_error(self._ctx).decode(errors="replace") if self._ctx else "allocation failure"
            self._lib.kpc_destroy(self._ctx)
            self._ctx = None
            raise KPopCountError(rc, msg)
        self._chunks = []
        self._sink = N.SINK_FN(self._on_text)
        self._check(self._lib.kpc_set_sink(self._ctx, self._sink, None))
        self._out_file = None

    # ---- plumbing -------------------------------------------------------------------------------------------
    def _on_text(self, _user, data, n):
        b = ctypes.string_at(data, n)
        if self._out_file is not None:
            self._out_file.write(b)
        else:
            self._chunks.append(b)
        return 0

    def _check(self, rc):
        if rc != N.KPC_OK:
            raise KPopCountError(rc, self._lib.kpc_error(self._ctx).decode(errors="replace"))

    def close(self):
        if self._ctx:
            self._lib.kpc_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- the C ABI, one method per entry point -----------------------------------------------------------------
    def begin(self, fmt):
        self._check(self._lib.kpc_begin(self._ctx, _FORMATS.get(fmt, fmt)))

    def feed(self, data, mate=0, eof=False):
        """data: bytes-like (host memory)."""
        buf = (ctypes.c_char * len(data)).from_buffer_copy(data) if len(data) else None
        self._check(self._lib.kpc_feed(self._ctx, mate, buf, len(data), 1 if eof else 0))

    def feed_pointer(self, ptr, n, mate=0, eof=False):
        """raw host pointer (e.g. a pinned torch tensor's data_ptr())."""
        self._check(self._lib.kpc_feed(self._ctx, mate, ctypes.c_void_p(ptr), n, 1 if eof else 0))

    def feed_device(self, ptr, n, mate=0, eof=True):
        """device pointer (16-byte aligned); the whole input in one call."""
        self._check(self._lib.kpc_feed_device(self._ctx, mate, ctypes.c_void_p(ptr), n, 1 if eof else 0))

    def set_pair_limit(self, n):
        self._check(self._lib.kpc_set_pair_limit(self._ctx, n))

    def set_single_pass(self, on=True):
        """weave paired files on the host in every mode: inputs are read once (pipes), a shorter mate ends the pair."""
        self._check(self._lib.kpc_set_single_pass(self._ctx, 1 if on else 0))

    def complete_pairs(self):
        return self._lib.kpc_complete_pairs(self._ctx)

    def end(self):
        self._check(self._lib.kpc_end(self._ctx))

    def finish(self):
        self._check(self._lib.kpc_finish(self._ctx))

    def reset(self):
        self._check(self._lib.kpc_reset(self._ctx))
        self._chunks = []

    def reset_label(self, label):
        """Empty the tables and give the next run this label: one context serves many samples (batch use)."""
        self._check(self._lib.kpc_reset_label(self._ctx, label.encode()))
        self.label = label
        self._chunks = []

    def discard_text(self, flag=True):
        self._check(self._lib.kpc_discard_text(self._ctx, 1 if flag else 0))

    def text_bytes(self):
        return self._lib.kpc_text_bytes(self._ctx)

    def sync(self):
        self._check(self._lib.kpc_sync(self._ctx))

    def stream_handle(self):
        return self._lib.kpc_stream(self._ctx)

    def kernel_launches(self):
        return self._lib.kpc_kernel_launches(self._ctx)

    def set_text_buffer(self, ptr, capacity):
        """Spectra text goes straight into the host buffer at `ptr` (e.g. a pinned tensor) instead of the callback."""
        self._check(self._lib.kpc_set_sink_buffer(self._ctx, ctypes.c_void_p(ptr), capacity))

    def text_buffer_used(self):
        return self._lib.kpc_sink_buffer_used(self._ctx)

    def profile_enable(self, on=True):
        """CUDA-event timing of the fast FASTQ kernels (measurement aid of bench.py)."""
        self._check(self._lib.kpc_profile_enable(self._ctx, 1 if on else 0))

    def profile_read(self):
        """(partition_ms, count_ms, launches, bytes) summed over the launches since the last read."""
        a, b = ctypes.c_double(), ctypes.c_double()
        n, nb = ctypes.c_ulonglong(), ctypes.c_ulonglong()
        self._check(self._lib.kpc_profile_read(self._ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n), ctypes.byref(nb)))
        return a.value, b.value, n.value, nb.value

    def kmers_counted(self):
        out = ctypes.c_ulonglong()
        self._check(self._lib.kpc_kmers_counted(self._ctx, ctypes.byref(out)))
        return out.value

    def dense_table(self):
        """(lo_ptr, hi_ptr or None, n_bins) of the device-resident 4^k table."""
        lo, hi, nb = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_ulonglong()
        self._check(self._lib.kpc_dense_table(self._ctx, ctypes.byref(lo), ctypes.byref(hi), ctypes.byref(nb)))
        return lo.value, hi.value, nb.value

    def dense_max(self):
        out = ctypes.c_ulonglong()
        self._check(self._lib.kpc_dense_max(self._ctx, ctypes.byref(out)))
        return out.value

    def set_record_base(self, first_record):
        self._check(self._lib.kpc_set_record_base(self._ctx, first_record))

    def hash_export(self):
        """(keys ptr, counts ptr, ranks ptr, slots) of the distinct k-mers counted so far (device arrays of u64)."""
        k, c, r = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        n = ctypes.c_ulonglong(0)
        self._check(self._lib.kpc_hash_export(self._ctx, ctypes.byref(k), ctypes.byref(c), ctypes.byref(r), ctypes.byref(n)))
        return k.value or 0, c.value or 0, r.value or 0, int(n.value)

    def hash_import(self, keys_ptr, counts_ptr, ranks_ptr, n, clear_first=True):
        self._check(self._lib.kpc_hash_import(self._ctx, ctypes.c_void_p(keys_ptr), ctypes.c_void_p(counts_ptr),
                                              ctypes.c_void_p(ranks_ptr), n, 1 if clear_first else 0))

    def bucket_count(self):
        return int(self._lib.kpc_bucket_count(self._ctx))

    def backend(self):
        return self._lib.kpc_backend().decode()

    def dense_has_hi(self):
        v = ctypes.c_int(0)
        self._check(self._lib.kpc_dense_has_hi(self._ctx, ctypes.byref(v)))
        return bool(v.value)

    def count_newlines(self, device_ptr, n):
        """Line feeds in n bytes of device memory (16-byte aligned pointer)."""
        v = ctypes.c_ulonglong(0)
        self._check(self._lib.kpc_count_newlines(self._ctx, ctypes.c_void_p(device_ptr), n, ctypes.byref(v)))
        return int(v.value)

    def dense_promote(self):
        self._check(self._lib.kpc_dense_promote(self._ctx))

    def synth_fastq(self, device_ptr, first_record, n_records, seed):
        self._check(self._lib.kpc_synth_fastq(self._ctx, ctypes.c_void_p(device_ptr), first_record, n_records, seed))

    def synth_offset(self, record):
        return self._lib.kpc_synth_offset(record)

    def take_text(self):
        out = b"".join(self._chunks)
        self._chunks = []
        return out

    # ---- KMerCounter.compute ----------------------------------------------------------------------------------
    def compute(self, inputs, fname="", chunk_bytes=32 << 20):
        """Count every input in order and dump the spectra; returns the text when fname == ''.

        Paired-end files of unequal length stop at the shorter one like FASTQ.iter_pe (Files.ml:228-247): the run is
        repeated with a pair limit, which is what KPC_E_PE_MISMATCH asks for.
        """
        limits = [-1] * len(inputs)
        while True:
            try:
                return self._compute_once(inputs, fname, chunk_bytes, limits)
            except KPopCountError as e:
                if e.code != N.KPC_E_PE_MISMATCH or limits[self._bad_input] >= 0:
                    raise
                limits[self._bad_input] = self.complete_pairs()
                self.reset()

    def _compute_once(self, inputs, fname, chunk_bytes, limits):
        has_pairs = any(i[0] == "paired-end" for i in inputs)
        self._chunks = []
        self._out_file = None
        if not inputs:
            return b"" if not fname else None  # bin/KPopCount.ml:218
        direct = bool(fname) and not has_pairs
        if direct:
            self._out_file = open(fname, "wb")
        try:
            for j, inp in enumerate(inputs):
                self._bad_input = j
                self.set_pair_limit(limits[j])
                self.begin(inp[0])
                paths = inp[1:]
                files = [open(p, "rb") for p in paths]
                try:
                    done = [False] * len(files)
                    while not all(done):
                        for m, f in enumerate(files):
                            if done[m]:
                                continue
                            data = f.read(chunk_bytes)
                            eof = len(data) < chunk_bytes
                            self.feed(data, mate=m, eof=eof)
                            done[m] = eof
                finally:
                    for f in files:
                        f.close()
                self.end()
            self.finish()
        finally:
            if self._out_file is not None:
                self._out_file.close()
                self._out_file = None
        text = self.take_text()
        if fname and not direct:
            with open(fname, "wb") as f:
                f.write(text)
            return None
        return None if direct else text
