(* Kpc_gpu.ml -- OCaml side of the C ABI of libkpopcount_gpu.so (include/kpopcount.h), one `external` per entry point
   the KPopCount host needs.  The C glue is kpc_stubs.c (same directory).  Every call raises
   [Failure (kpc_error ctx)] when the library returns a negative code, except that argument errors raise
   [Invalid_argument]: left uncaught in bin/KPopCount.ml they end the process with exit code 2, exactly like the
   [Failure "...Malformed FASTQ..."] of BiOCamLib/lib/Files.ml:213-214 and [Quotes_in_name]
   (BiOCamLib/lib/Matrix.ml:83-99) do in the reference.

   This file replaces nothing in the reference by itself: it is what the new [KMerCounter.compute]
   (KPopCount_gpu.ml, replacing bin/KPopCount.ml:20-64) is written against. *)

type ctx
(* custom block around kpc_ctx*; finalised with kpc_destroy *)

type content = DNA_ss | DNA_ds | Protein
(* = KPC_DNA_SS | KPC_DNA_DS | KPC_PROTEIN: the order of Content.t, bin/KPopCount.ml:66-70 *)

type format = FASTA | FASTQ_SE | FASTQ_PE
(* = KPC_FASTA | KPC_FASTQ_SE | KPC_FASTQ_PE: the constructors of Files.Type.t that KPopCount accepts,
   BiOCamLib/lib/Files.ml:315-323 *)

type pinned = (char, Bigarray.int8_unsigned_elt, Bigarray.c_layout) Bigarray.Array1.t
(* a view (CAML_BA_EXTERNAL) of one of the context's pinned staging buffers *)

(* KIHF.create max_results_size + the functor's k check (bin/KPopCount.ml:35, 239-249).
   [devices] lists the CUDA devices to use ([| 0 |] for one GPU); label "" selects -L behaviour. *)
external create : k:int -> content -> max_results_size:int -> label:string -> devices:int array -> ctx
  = "kpc_ml_create"
(* stdout / open_out fname (bin/KPopCount.ml:27-31): the spectra text is written to this descriptor, in final order *)
external set_out : ctx -> Unix.file_descr -> unit = "kpc_ml_set_out"
external staging_slots : ctx -> int = "kpc_ml_staging_slots"
external staging : ctx -> int -> pinned = "kpc_ml_staging"
(* one element of Files.ReadsIterate.t (Files.ml:348-349); the first call writes the "\t<label>\n" header *)
external begin_ : ctx -> format -> unit = "kpc_ml_begin"
(* raw file bytes: replaces input_line, the linter, KIH.iterc and KIHF.add (Files.ml:96-122, 201-250;
   Sequences.ml:41-67; KMers.ml:357-389, 107-111).  Releases the runtime lock while it blocks. *)
external feed : ctx -> mate:int -> pinned -> len:int -> eof:bool -> unit = "kpc_ml_feed"
(* the same for bytes in an ordinary string / Bytes.t (copied before the call returns) *)
external feed_bytes : ctx -> mate:int -> Bytes.t -> len:int -> eof:bool -> unit = "kpc_ml_feed_bytes"
external end_ : ctx -> unit = "kpc_ml_end"
(* the final KIHF.iter (bin/KPopCount.ml:60) *)
external finish : ctx -> unit = "kpc_ml_finish"
(* FASTQ.iter_pe stops at the shorter file (Files.ml:228-247) *)
external set_pair_limit : ctx -> int -> unit = "kpc_ml_set_pair_limit"
external complete_pairs : ctx -> int = "kpc_ml_complete_pairs"
(* one pass over paired files in every mode (the reference reads its inputs once, and they may be pipes) *)
external set_single_pass : ctx -> bool -> unit = "kpc_ml_set_single_pass"
external kmers_counted : ctx -> int = "kpc_ml_kmers_counted"
external reset : ctx -> unit = "kpc_ml_reset"
external backend : unit -> string = "kpc_ml_backend"
