(* KPopCount_gpu.ml -- the GPU-backed KMerCounter for bin/KPopCount.ml.

   What changes in the reference's bin/KPopCount.ml (v18, a1fda68):
     * lines 20-64 (functor KMerCounter: ReadsIterate.iter + KIH.iterc + KIHF table + the two KIHF.iter dumps) are
       replaced by the module below: the device does record splitting, linting, k-mer construction, counting, the
       -L / -M dump rule and the text formatting; the host only streams raw bytes;
     * lines 239-249 (functor applications per content type) become the single call at the bottom of this file;
     * everything else -- Content, Parameters, the Tools.Argv table (lines 66-212), the -l/-L check (213-214), the
       FASTA/FASTQ mixing check (220-236), Spectra.make_filename (185) -- stays as it is.
   Argument meaning and error behaviour are those of the reference: header first (bin/KPopCount.ml:33-34), inputs in
   argv order (217), exceptions left uncaught (exit 2). *)

module KMerCounter :
  sig
    val compute: ?verbose:bool -> content:Kpc_gpu.content -> k:int -> Files.Type.t list -> int -> string -> string -> unit
  end
= struct
    let compute ?(verbose = false) ~content ~k inputs max_results_size label fname =
      let fd =
        if fname = "" then
          Unix.stdout
        else
          Unix.openfile fname [ Unix.O_WRONLY; Unix.O_CREAT; Unix.O_TRUNC ] 0o644 in
      (* KPC_DEVICES=0,1,... selects the GPUs (argv must stay the reference's); default: device 0 *)
      let devices =
        match Sys.getenv_opt "KPC_DEVICES" with
        | None | Some "" -> [| 0 |]
        | Some s -> String.split_on_char ',' s |> List.map int_of_string |> Array.of_list in
      let ctx = Kpc_gpu.create ~k content ~max_results_size ~label ~devices in
      (* The "\t<label>\n" header is written by the library when the first input begins *)
      Kpc_gpu.set_out ctx fd;
      let n_slots = Kpc_gpu.staging_slots ctx and slot = ref 0 and reads_bytes = ref 0 in
      (* Raw bytes, no input_line, no linter: both happen on the device.
         Works on pipes (/dev/stdin) as the reference does: nothing is seeked or mapped *)
      let scratch = Bytes.create 65536 in
      (* one staging buffer's worth of [ic] to the device; true at end of file *)
      let feed_some mate ic =
        let buf = Kpc_gpu.staging ctx !slot in
        let cap = Bigarray.Array1.dim buf in
        (* Fill the pinned buffer (Unix.read_bigarray needs OCaml >= 5.2; this works from 4.12 on) *)
        let filled = ref 0 and eof = ref false in
        while not !eof && !filled < cap do
          let want = min (Bytes.length scratch) (cap - !filled) in
          let n = Unix.read ic scratch 0 want in
          if n = 0 then
            eof := true
          else begin
            for i = 0 to n - 1 do
              Bigarray.Array1.unsafe_set buf (!filled + i) (Bytes.unsafe_get scratch i)
            done;
            filled := !filled + n
          end
        done;
        Kpc_gpu.feed ctx ~mate buf ~len:!filled ~eof:!eof;
        reads_bytes := !reads_bytes + !filled;
        if verbose then
          Printf.eprintf "%s\r(%s): Streamed %d bytes%!" String.TermIO.clear __FUNCTION__ !reads_bytes;
        slot := (!slot + 1) mod n_slots;
        !eof in
      let stream path =
        let ic = Unix.openfile path [ Unix.O_RDONLY ] 0 in
        while not (feed_some 0 ic) do () done;
        Unix.close ic
      (* the two files of a pair advance together, one buffer each in turn, as FASTQ.iter_pe reads them *)
      and stream_pair path1 path2 =
        let ic1 = Unix.openfile path1 [ Unix.O_RDONLY ] 0 in
        let ic2 = Unix.openfile path2 [ Unix.O_RDONLY ] 0 in
        let eof1 = ref false and eof2 = ref false in
        while not (!eof1 && !eof2) do
          if not !eof1 then eof1 := feed_some 0 ic1;
          if not !eof2 then eof2 := feed_some 1 ic2
        done;
        Unix.close ic1;
        Unix.close ic2 in
      (* inputs are read once, like the reference does: a shorter mate file ends its pair of files (Files.ml:228-247) *)
      Kpc_gpu.set_single_pass ctx true;
      List.iter
        (function
          | Files.Type.FASTA file ->
            Kpc_gpu.begin_ ctx Kpc_gpu.FASTA; stream file; Kpc_gpu.end_ ctx
          | SingleEndFASTQ file ->
            Kpc_gpu.begin_ ctx Kpc_gpu.FASTQ_SE; stream file; Kpc_gpu.end_ ctx
          | PairedEndFASTQ (file1, file2) ->
            (* FASTQ.iter_pe alternates the mates pair by pair (Files.ml:222-250); the library re-creates that order *)
            Kpc_gpu.begin_ ctx Kpc_gpu.FASTQ_PE; stream_pair file1 file2; Kpc_gpu.end_ ctx
          | InterleavedFASTQ _ | Tabular _ ->
            assert false) (* not reachable from KPopCount's argv, bin/KPopCount.ml:140,147,157 *)
        inputs;
      if verbose then
        Printf.eprintf "%s\r(%s): Streamed %d bytes.\n(%s): Outputting hashes...%!"
          String.TermIO.clear __FUNCTION__ !reads_bytes __FUNCTION__;
      (* The final KIHF.iter, bin/KPopCount.ml:60 *)
      Kpc_gpu.finish ctx;
      if verbose then
        Printf.eprintf " done.%!\n";
      if fname <> "" then
        Unix.close fd
  end

(* Replacement for bin/KPopCount.ml:239-249 -- inside `if !Parameters.inputs <> [] then begin ... end`, after the
   FASTA/FASTQ mixing check, which keeps iterating over !Parameters.inputs (a Files.Type.t list):

    KMerCounter.compute ~verbose:!Parameters.verbose
      ~content:(match !Parameters.content with
                | DNA_ss -> Kpc_gpu.DNA_ss | DNA_ds -> Kpc_gpu.DNA_ds | Protein -> Kpc_gpu.Protein)
      ~k:!Parameters.k !Parameters.inputs !Parameters.max_results_size !Parameters.label !Parameters.output
*)
