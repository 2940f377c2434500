(*---------------------------------------------------------------------------
   zipc_cuda -- the zipc hot path on NVIDIA B200, behind zipc's own signatures.

   NOT COMPILED IN THIS REPOSITORY'S IMAGE (there is no OCaml toolchain there);
   kept minimal and mechanical.  See INTEGRATION.md.
  ---------------------------------------------------------------------------*)

(** [Zipc_cuda] satisfies [module type of Zipc_deflate] for everything on the
    hot path, so that [module Zipc_deflate = Zipc_cuda] type-checks in a client,
    and adds the batch entry points a GPU needs. *)

type uint16 = int
type uint32 = int32

module Crc_32 : sig
  type t = uint32
  val equal : t -> t -> bool
  val check : expect:t -> found:t -> (unit, string) result
  val pp : Format.formatter -> t -> unit
  val string : ?start:int -> ?len:int -> string -> t
  val strings : string array -> t array
  (** One GPU call for all strings. *)
end

module Adler_32 : sig
  type t = uint32
  val equal : t -> t -> bool
  val check : expect:t -> found:t -> (unit, string) result
  val pp : Format.formatter -> t -> unit
  val string : ?start:int -> ?len:int -> string -> t
  (** Bit-exact with [Zipc_deflate.Adler_32.string] (signed remainder included). *)
end

val inflate :
  ?decompressed_size:int -> ?start:int -> ?len:int -> string -> (string, string) result
val inflate_and_crc_32 :
  ?decompressed_size:int -> ?start:int -> ?len:int -> string -> (string * Crc_32.t, string) result
val inflate_and_adler_32 :
  ?decompressed_size:int -> ?start:int -> ?len:int -> string -> (string * Adler_32.t, string) result
val zlib_decompress :
  ?decompressed_size:int -> ?start:int -> ?len:int -> string ->
  (string * Adler_32.t, (Adler_32.t * Adler_32.t) option * string) result

type level = [ `None | `Fast | `Default | `Best ]

val deflate : ?level:level -> ?start:int -> ?len:int -> string -> (string, string) result
val crc_32_and_deflate :
  ?level:level -> ?start:int -> ?len:int -> string -> (Crc_32.t * string, string) result
val adler_32_and_deflate :
  ?level:level -> ?start:int -> ?len:int -> string -> (Adler_32.t * string, string) result
val zlib_compress :
  ?level:level -> ?start:int -> ?len:int -> string -> (Adler_32.t * string, string) result

(** {1 Batch forms} one call for n members; results in input order. *)

type crc_op = Nop | Adler_32_op | Crc_32_op

val inflate_batch :
  crc_op:crc_op -> (string * int * int * int option) array ->
  (string * uint32, string) result array
(** [(s, start, len, decompressed_size)] per member. *)

val deflate_batch :
  ?level:level -> crc_op:crc_op -> string array -> (uint32 * string, string) result array

(** {1 One large stream as independent segments}

    [deflate_segmented s] compresses [s] as window-reset segments of
    [segment_size] bytes (default 256 KiB), one CTA each, joined byte aligned
    into ONE valid RFC 1951 stream that [Zipc_deflate.inflate] reads; the
    index holds [(compressed offset, uncompressed offset)] per segment plus
    the totals.  [inflate_segmented ~index] decodes one segment per warp. *)

val deflate_segmented :
  ?level:level -> ?segment_size:int -> string ->
  (string * (int * int) array * Crc_32.t, string) result
val inflate_segmented :
  index:(int * int) array -> string -> (string * Crc_32.t, string) result

(** {1 ZIP layer} batch forms of [Zipc.File] and [Zipc.to_binary_string]
    (the documented plug point, zipc.mli:26-28,100-121). *)

module File : sig
  val deflate_of_binary_strings :
    ?level:level -> string array -> (Zipc.File.t, string) result array
  (** [Zipc.File.deflate_of_binary_string] for every payload, one GPU call. *)

  val to_binary_strings : Zipc.File.t array -> (string, string) result array
  (** [Zipc.File.to_binary_string] for every file (CRC-32 checked), one GPU
      call; same error strings, ["deflate: "] prefix included. *)
end

val archive_to_binary_string :
  ?level:level -> ?first:Zipc.Fpath.t -> (Zipc.Fpath.t * string) array ->
  (string, string) result
(** [File.deflate_of_binary_string] of every payload, [Member.make] with the
    default mode and mtime, [Zipc.add] and [Zipc.to_binary_string ?first] in
    one call: payloads are compressed and gathered to their archive offsets
    on the GPU. *)

(** {1 Context} *)

val set_device : int -> unit
(** CUDA device used by this process (default 0).  One context per device. *)

val set_devices : int -> unit
(** Box-wide mode: bit [d] of the mask selects CUDA device [d], [0] = all.
    The batch entry points then partition their members over these devices
    inside one call (zipc_b200_multi_*).  [set_device] switches back. *)
