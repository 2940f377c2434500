(*---------------------------------------------------------------------------
   zipc_cuda -- OCaml side of the binding.  Thin: argument defaults, result
   mapping and error strings; all work happens in libzipc_b200.so.
   NOT COMPILED IN THIS REPOSITORY'S IMAGE (no OCaml toolchain); see INTEGRATION.md.
  ---------------------------------------------------------------------------*)

type uint16 = int
type uint32 = int32
type level = [ `None | `Fast | `Default | `Best ]
type crc_op = Nop | Adler_32_op | Crc_32_op

(* status codes of include/zipc_b200.h *)
let ok = 0 and err_checksum = 6 and err_zlib_method = 3 and err_invalid_arg = 8

external strerror : int -> string = "zipc_cuda_strerror"
external set_device : int -> unit = "zipc_cuda_set_device"
external set_devices : int -> unit = "zipc_cuda_set_devices"

(* Each stub copies its inputs into pinned staging memory BEFORE releasing the
   runtime lock (the GC may move strings while it is released) and allocates the
   result strings after re-acquiring it. *)
external crc32_batch_stub : string array -> int array -> int array -> int32 array
  = "zipc_cuda_crc32_batch"
external adler32_stub : string -> int -> int -> int32 = "zipc_cuda_adler32"
external inflate_batch_stub :
  int (* crc_op *) -> string array -> int array -> int array -> int array (* -1 = unknown *) ->
  (int * string * int32) array = "zipc_cuda_inflate_batch"
external deflate_batch_stub :
  int (* level *) -> int (* crc_op *) -> string array -> int array -> int array ->
  (int * string * int32) array = "zipc_cuda_deflate_batch"
external zlib_decompress_stub :
  string -> int -> int -> int -> int * string * int32 * int32 = "zipc_cuda_zlib_decompress"
external zlib_compress_stub :
  int -> string -> int -> int -> int * string * int32 = "zipc_cuda_zlib_compress"
external deflate_segmented_stub :
  int (* level *) -> string -> int (* segment_size *) -> int * string * int array * int32
  = "zipc_cuda_deflate_segmented"
external inflate_segmented_stub :
  string -> int array (* flat index *) -> int * string * int32 = "zipc_cuda_inflate_segmented"
external archive_stub :
  int (* level *) -> string array (* paths *) -> string array (* payloads *) -> string (* first, "" = default *) ->
  int * string = "zipc_cuda_archive"

let range ?(start = 0) ?len s =
  let len = match len with None -> String.length s - start | Some l -> l in
  if start < 0 || len < 0 || start + len > String.length s
  then invalid_arg "index out of bounds" else (start, len)

let crc_error e f =
  Error (Printf.sprintf "Checksum mismatch, expected %lx found %lx)" e f)

module Crc_32 = struct
  type t = uint32
  let equal = Int32.equal
  let pp ppf c = Format.fprintf ppf "%lx" c
  let check ~expect:e ~found:f = if equal e f then Ok () else crc_error e f
  let strings ss =
    crc32_batch_stub ss (Array.map (fun _ -> 0) ss) (Array.map String.length ss)
  let string ?start ?len s =
    let start, len = range ?start ?len s in
    (crc32_batch_stub [| s |] [| start |] [| len |]).(0)
end

module Adler_32 = struct
  type t = uint32
  let equal = Int32.equal
  let pp ppf c = Format.fprintf ppf "%lx" c
  let check ~expect:e ~found:f = if equal e f then Ok () else crc_error e f
  let string ?start ?len s =
    let start, len = range ?start ?len s in adler32_stub s start len
end

let int_of_crc_op = function Nop -> 0 | Adler_32_op -> 1 | Crc_32_op -> 2
let int_of_level = function `None -> 0 | `Fast -> 1 | `Default -> 2 | `Best -> 3

let result_of (st, s, c) = if st = ok then Ok (s, c) else Error (strerror st)

let inflate_batch ~crc_op ms =
  let ss = Array.map (fun (s, _, _, _) -> s) ms in
  let starts = Array.map (fun (_, st, _, _) -> st) ms in
  let lens = Array.map (fun (_, _, l, _) -> l) ms in
  let dsz = Array.map (fun (_, _, _, d) -> Option.value d ~default:(-1)) ms in
  Array.map result_of (inflate_batch_stub (int_of_crc_op crc_op) ss starts lens dsz)

let inflate1 crc_op ?decompressed_size ?start ?len s =
  let start, len = range ?start ?len s in
  (inflate_batch ~crc_op [| (s, start, len, decompressed_size) |]).(0)

let inflate ?decompressed_size ?start ?len s =
  Result.map fst (inflate1 Nop ?decompressed_size ?start ?len s)
let inflate_and_crc_32 ?decompressed_size ?start ?len s =
  inflate1 Crc_32_op ?decompressed_size ?start ?len s
let inflate_and_adler_32 ?decompressed_size ?start ?len s =
  inflate1 Adler_32_op ?decompressed_size ?start ?len s

let zlib_decompress ?decompressed_size ?start ?len s =
  let start, len = range ?start ?len s in
  let d = Option.value decompressed_size ~default:(-1) in
  match zlib_decompress_stub s start len d with
  | st, out, _, found when st = ok -> Ok (out, found)
  | st, _, expect, found when st = err_checksum ->
      Error (Some (expect, found),
             Printf.sprintf "Checksum mismatch, expected %lx found %lx)" expect found)
  | st, _, _, cm when st = err_zlib_method ->
      Error (None, Printf.sprintf "Unknown compression method (%ld)" cm)
  | st, _, _, _ -> Error (None, strerror st)

let deflate_batch ?(level = `Default) ~crc_op ss =
  let r = deflate_batch_stub (int_of_level level) (int_of_crc_op crc_op) ss
      (Array.map (fun _ -> 0) ss) (Array.map String.length ss) in
  Array.map (fun (st, s, c) -> if st = ok then Ok (c, s) else Error (strerror st)) r

let deflate1 crc_op ?(level = `Default) ?start ?len s =
  let start, len = range ?start ?len s in
  match deflate_batch_stub (int_of_level level) (int_of_crc_op crc_op) [| s |] [| start |] [| len |] with
  | [| (st, cs, c) |] when st = ok -> Ok (c, cs)
  | [| (st, _, _) |] -> Error (strerror st)
  | _ -> assert false

let deflate ?level ?start ?len s = Result.map snd (deflate1 Nop ?level ?start ?len s)
let crc_32_and_deflate ?level ?start ?len s = deflate1 Crc_32_op ?level ?start ?len s
let adler_32_and_deflate ?level ?start ?len s = deflate1 Adler_32_op ?level ?start ?len s

let zlib_compress ?(level = `Default) ?start ?len s =
  let start, len = range ?start ?len s in
  match zlib_compress_stub (int_of_level level) s start len with
  | st, zs, adler when st = ok -> Ok (adler, zs)
  | st, _, _ -> Error (strerror st)

(* ---- one large stream as independent segments ---- *)

let deflate_segmented ?(level = `Default) ?(segment_size = 256 * 1024) s =
  match deflate_segmented_stub (int_of_level level) s segment_size with
  | st, cs, flat, crc when st = ok ->
      Ok (cs, Array.init (Array.length flat / 2) (fun i -> (flat.(2 * i), flat.(2 * i + 1))), crc)
  | st, _, _, _ -> Error (strerror st)

let inflate_segmented ~index s =
  let flat = Array.make (2 * Array.length index) 0 in
  Array.iteri (fun i (c, u) -> flat.(2 * i) <- c; flat.(2 * i + 1) <- u) index;
  match inflate_segmented_stub s flat with
  | st, out, crc when st = ok -> Ok (out, crc)
  | st, _, _ -> Error (strerror st)

(* ---- ZIP layer ---- *)

module File = struct
  let deflate_of_binary_strings ?level ss =
    Array.map2 (fun s -> function
      | Error _ as e -> e
      | Ok (crc, cs) ->
          Zipc.File.make ~compression:Zipc.Deflate cs
            ~decompressed_size:(String.length s) ~decompressed_crc_32:crc)
      ss (deflate_batch ?level ~crc_op:Crc_32_op ss)

  (* zipc.ml:205-225: encrypted -> error; Stored / Deflate handled; anything else -> error; then the CRC check *)
  let to_binary_strings fs =
    let job f =
      if Zipc.File.is_encrypted f then `Err "Encrypted files are not supported" else
      match Zipc.File.compression f with
      | Zipc.Stored | Zipc.Deflate -> `Gpu
      | c -> `Err (Format.asprintf "Compression %a not supported" Zipc.pp_compression c)
    in
    let jobs = Array.map job fs in
    let idx = List.filter (fun i -> jobs.(i) = `Gpu) (List.init (Array.length fs) Fun.id) in
    let deflated = List.filter (fun i -> Zipc.File.compression fs.(i) = Zipc.Deflate) idx in
    let stored = List.filter (fun i -> Zipc.File.compression fs.(i) = Zipc.Stored) idx in
    let arg i =
      let f = fs.(i) in
      (Zipc.File.compressed_bytes f, Zipc.File.start f, Zipc.File.compressed_size f,
       Some (Zipc.File.decompressed_size f))
    in
    let inflated = inflate_batch ~crc_op:Crc_32_op (Array.of_list (List.map arg deflated)) in
    let stored_crcs =
      let a = Array.of_list stored in
      crc32_batch_stub (Array.map (fun i -> Zipc.File.compressed_bytes fs.(i)) a)
        (Array.map (fun i -> Zipc.File.start fs.(i)) a)
        (Array.map (fun i -> Zipc.File.compressed_size fs.(i)) a)
    in
    let out = Array.map (function `Err m -> Error m | `Gpu -> Error "") jobs in
    let check i s found =
      let expect = Zipc.File.decompressed_crc_32 fs.(i) in
      out.(i) <- (match Crc_32.check ~expect ~found with Ok () -> Ok s | Error _ as e -> e)
    in
    List.iteri (fun k i -> match inflated.(k) with
      | Ok (s, crc) -> check i s crc
      | Error m -> out.(i) <- Error ("deflate: " ^ m)) deflated;
    List.iteri (fun k i ->
      let f = fs.(i) in
      check i (String.sub (Zipc.File.compressed_bytes f) (Zipc.File.start f) (Zipc.File.compressed_size f))
        stored_crcs.(k)) stored;
    out
end

let archive_to_binary_string ?(level = `Default) ?(first = "") members =
  match archive_stub (int_of_level level) (Array.map fst members) (Array.map snd members) first with
  | st, archive when st = ok -> Ok archive
  | st, _ -> Error (strerror st)
