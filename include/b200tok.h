/*
 * b200tok.h — C ABI of the B200-native tokenizer hot path (libb200tok.so).
 *
 * This is the boundary an OpenVINO-side `ov::Op::evaluate()` (or any other host) binds to.  Each
 * entry point replaces one reference `evaluate()` body and takes the reference's own decomposed
 * tensors — `(begins:i32, ends:i32, chars:u8)` strings, with `(ragged_begins:i32[B], ragged_ends:i32[B])`
 * in front for ragged tensors (reference src/utils.cpp:84-102) — as plain pointers + sizes.
 *
 *   b200tok_regexsplit_*   replaces RegexSplit::evaluate          src/regex_split.cpp:124-324
 *   b200tok_specialsplit_* replaces SpecialTokensSplit::evaluate  src/special_tokens_split.cpp:61-162
 *   b200tok_bpe_*          replaces BPETokenizer::evaluate        src/bpe_tokenizer.cpp:47-164
 *                          (+ BPETokenizerImpl ctor :341-388, tokenize_into :196-339)
 *   b200tok_wordpiece_*    replaces WordpieceTokenizer::evaluate  src/wordpiece_tokenizer.cpp:49-133
 *   b200tok_vocabenc_*     replaces VocabEncoder::evaluate_impl   src/vocab_encoder.cpp:56-94
 *   b200tok_vocabdec_*     replaces VocabDecoder::evaluate        src/vocab_decoder.cpp:23-87
 *   b200tok_bytefallback_run replaces ByteFallback::evaluate      src/byte_fallback.cpp:16-50
 *   b200tok_truncate_run / b200tok_combine_segments_run / b200tok_ragged_to_dense_run replace Truncate / CombineSegments /
 *                          RaggedToDense::evaluate (src/truncate.cpp:37-147, src/combine_segments.cpp:36-134,
 *                          src/ragged_to_dense.cpp:70-174); b200tok_post_dense_run fuses the three
 *   b200tok_split_bpe_run / b200tok_split_wordpiece_run  fuse RegexSplit(+RegexSplit) -> tokenizer
 *                          in one kernel (pieces never leave shared memory); same results as the
 *                          two ops run back to back.
 *
 * Conventions
 *   - Every function returns 0 on success or a negative B200TOK_E_* code; the message is available
 *     from b200tok_last_error() (thread-local).  Nothing throws across this boundary; the ov::Op
 *     shim rethrows as ov::Exception (reference error convention: OPENVINO_ASSERT/THROW).
 *   - `*_create` corresponds to the reference's lazy first-call initialisation (std::call_once in
 *     evaluate): it reads the Constant inputs once and builds device-resident tables.
 *   - Handles may be used from several threads; `*_run` calls on one handle are serialised inside.
 *   - `mem` says where the data pointers of a run live: B200TOK_MEM_HOST (pageable or pinned host
 *     memory, as ov::Tensor::data() gives; the library stages H2D/D2H itself and returns when the
 *     outputs are complete) or B200TOK_MEM_DEVICE (device pointers on the handle's GPU; work is
 *     enqueued on `stream`).
 *   - Output buffers are caller-allocated at the reference's own worst-case sizes (noted per op) and
 *     the number of produced elements is returned, mirroring `set_shape(worst) ... set_shape(actual)`.
 *   - There is no CPU fallback: without a usable sm_100 device `*_create` fails with B200TOK_E_CUDA.
 */
#ifndef B200TOK_H_
#define B200TOK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200TOK_API __attribute__((visibility("default")))

#define B200TOK_OK 0
#define B200TOK_E_INVALID (-1)     /* bad argument / malformed tensors                                 */
#define B200TOK_E_CUDA (-2)        /* CUDA runtime error or no device                                   */
#define B200TOK_E_CAPACITY (-3)    /* caller's output buffer is smaller than the result                 */
#define B200TOK_E_UNSUPPORTED (-4) /* e.g. a split pattern the GPU splitter does not recognise          */
#define B200TOK_E_VOCAB (-5)       /* merge refers to a token missing from the vocab (ref: vocab.at())  */

#define B200TOK_MEM_HOST 0
#define B200TOK_MEM_DEVICE 1

typedef struct b200tok_object* b200tok_handle;

/* A decomposed string tensor (host memory; used for Constant inputs at create time). */
typedef struct {
    const int32_t* begins;
    const int32_t* ends;
    const uint8_t* chars;
    int64_t n;        /* number of strings  */
    int64_t n_chars;  /* length of `chars`  */
} b200tok_strings;

/* A ragged string tensor = inputs [0..4] of RegexSplit / BPETokenizer / WordpieceTokenizer
 * (+ optional input [5] `skips` of the 7-input RegexSplit form, src/regex_split.cpp:193-199). */
typedef struct {
    const int32_t* ragged_begins;  /* [n_rows]  */
    const int32_t* ragged_ends;    /* [n_rows]  */
    int64_t n_rows;
    const int32_t* begins;         /* [n_elems] */
    const int32_t* ends;           /* [n_elems] */
    int64_t n_elems;
    const uint8_t* chars;          /* [n_chars] */
    int64_t n_chars;
    const uint8_t* skips;          /* [n_elems] bool, or NULL */
    int mem;                       /* B200TOK_MEM_*  */
} b200tok_ragged_strings;

/* Ragged string result (RegexSplit outputs [0..3] and [5]; `chars` is aliased by the op,
 * src/regex_split.cpp:203).  Worst case n_elems + n_chars elements (src/regex_split.cpp:182). */
typedef struct {
    int32_t* ragged_begins;  /* [n_rows]   */
    int32_t* ragged_ends;    /* [n_rows]   */
    int32_t* begins;         /* [capacity] */
    int32_t* ends;           /* [capacity] */
    uint8_t* skips;          /* [capacity] or NULL */
    int64_t capacity;
    int64_t n_elems;         /* out: produced elements */
    int64_t n_rows;          /* out: rows written (1 for the whole-batch-empty shortcut, src/regex_split.cpp:129-143) */
    int mem;
} b200tok_ragged_strings_out;

/* Ragged i32 result (BPETokenizer / WordpieceTokenizer outputs [0..2]).  Worst case n_chars ids
 * (src/bpe_tokenizer.cpp:135, src/wordpiece_tokenizer.cpp:86). */
typedef struct {
    int32_t* begins;    /* [n_rows]   */
    int32_t* ends;      /* [n_rows]   */
    int32_t* ids;       /* [capacity] */
    int64_t capacity;
    int64_t n_ids;      /* out: produced ids (host value; valid on return unless n_ids_device is used) */
    int64_t* n_ids_device; /* MEM_DEVICE only, optional: if non-NULL the count is written here on the device
                              and the call does not synchronise the stream (fully asynchronous)        */
    int mem;
} b200tok_ragged_ids;

/* ------------------------------------------------------------------------------------------ */
B200TOK_API int b200tok_version(void);
B200TOK_API const char* b200tok_last_error(void);
B200TOK_API int b200tok_device_count(void);
B200TOK_API void b200tok_destroy(b200tok_handle h);     /* any handle kind */
/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
B200TOK_API int64_t b200tok_launch_count(b200tok_handle h);
/* Optional CUDA-event timing of the dominant kernel (the row kernel) of the row ops: enable once, then after a
 * call has completed (stream synchronised) read the duration of that call's row kernel in milliseconds.
 * Returns a negative value if timing is disabled or no call has been made. */
B200TOK_API void b200tok_set_timing(b200tok_handle h, int enabled);
B200TOK_API float b200tok_last_kernel_ms(b200tok_handle h);

/* ---- RegexSplit ----------------------------------------------------------------------------
 * attributes: behaviour / invert / max_splits (src/regex_split.hpp:42-47); pattern = input [5|6].
 * The GPU splitter hand-codes the tokenizer patterns the reference converter emits (GPT-2
 * byte-level, its individual-digits variant, Llama-3/cl100k, BERT whitespace, BERT punctuation,
 * `\w+|[^\w\s]+`, single-literal metaspace, `.`); any other pattern => B200TOK_E_UNSUPPORTED. */
typedef struct {
    const char* pattern;
    int64_t pattern_len;
    const char* behaviour;   /* "remove" | "isolate" | "contiguous" | "mergedwithprevious" | "mergedwithnext" */
    int invert;
    int max_splits;          /* -1 or > 0 */
    int device;              /* CUDA device ordinal */
} b200tok_regexsplit_desc;
B200TOK_API int b200tok_regexsplit_create(const b200tok_regexsplit_desc* desc, b200tok_handle* out);
B200TOK_API int b200tok_regexsplit_run(b200tok_handle h, const b200tok_ragged_strings* in,
                                       b200tok_ragged_strings_out* out, void* cuda_stream);
/* Legacy 9-input form (src/regex_split.cpp:102-113, 164-178, 231-238: inputs [6..8] = skip-token strings): elements equal to one of
 * `tokens` pass through unsplit (no skip flag is produced; the form has 5 outputs).  Call once after create, like the reference builds
 * its set on the first evaluate(); an empty list leaves the set unset.  Excludes the `skips` tensor of the 7-input form. */
B200TOK_API int b200tok_regexsplit_set_skip_tokens(b200tok_handle h, const b200tok_strings* tokens);

/* ---- SpecialTokensSplit ---------------------------------------------------------------------
 * replaces SpecialTokensSplit::evaluate (src/special_tokens_split.cpp:61-162); pattern = input [5|6], the alternation of
 *   (?:\s*)?(tok|tok|...)(?:\s*)?   groups the converter builds (python/openvino_tokenizers/tokenizer_pipeline.py:138-158);
 * anything else => B200TOK_E_UNSUPPORTED.  in->skips = input [5] of the 7-input form (may be NULL); outputs as RegexSplit's,
 * with out->skips = output [5] (required).  Worst case n_chars + n_elems elements.                                    */
B200TOK_API int b200tok_specialsplit_create(const char* pattern, int64_t pattern_len, int device, b200tok_handle* out);
B200TOK_API int b200tok_specialsplit_run(b200tok_handle h, const b200tok_ragged_strings* in,
                                         b200tok_ragged_strings_out* out, void* cuda_stream);

/* ---- BPETokenizer --------------------------------------------------------------------------
 * Constant inputs [5..] of the 11/14/15/18-input forms (src/bpe_tokenizer.cpp:18-21,69-114):
 * merges_right.begins == NULL selects the "L R" string form (11/15 inputs); added.n == 0 means
 * no added-token inputs.  Attributes as in src/bpe_tokenizer.hpp:220-228. */
typedef struct {
    b200tok_strings vocab;
    b200tok_strings merges_left;
    b200tok_strings merges_right;
    b200tok_strings added_tokens;
    const int32_t* added_ids;
    const char* unk_token;        int64_t unk_token_len;
    const char* suffix_indicator; int64_t suffix_indicator_len;   /* stored, unused (as in the reference) */
    const char* end_suffix;       int64_t end_suffix_len;
    int fuse_unk;
    int byte_fallback;
    int64_t cache_capacity;       /* accepted for IR compatibility; the GPU path has no result cache */
    int device;
} b200tok_bpe_desc;
B200TOK_API int b200tok_bpe_create(const b200tok_bpe_desc* desc, b200tok_handle* out);
B200TOK_API int b200tok_bpe_run(b200tok_handle h, const b200tok_ragged_strings* in, b200tok_ragged_ids* out,
                                void* cuda_stream);
/* Fused RegexSplit -> BPETokenizer (what the converted gpt2 / Llama-3 IRs chain back to back). */
B200TOK_API int b200tok_split_bpe_run(b200tok_handle split, b200tok_handle bpe, const b200tok_ragged_strings* in,
                                      b200tok_ragged_ids* out, void* cuda_stream);

/* ---- Sharded output over NVLink peer memory --------------------------------------------------------
 * One process per GPU tokenises its own row shard; instead of "emit locally, then all-gatherv", the emit step of this call
 * stores every id row straight into the result buffers of ALL ranks (peer-mapped device pointers, e.g. from CUDA IPC or
 * torch symmetric memory), so the exchange is fused into the compaction kernel.  Rank r's rows land in slot r:
 *   ids[p][r * slot_capacity + ...], begins/ends[p][r * rows_per_rank + row] (offsets already shifted by r * slot_capacity).
 * The caller orders a cross-rank barrier on the stream before reading (peers finish their stores at their own pace).
 * `in` must be device memory; every rank must pass the same world, slot_capacity (>= in->n_chars + in->n_elems * suffix)
 * and rows_per_rank (>= in->n_rows).                                                                              */
#define B200TOK_MAX_PEERS 8
typedef struct {
    int world, rank;
    int32_t* ids[B200TOK_MAX_PEERS];      /* each [world * slot_capacity]  */
    int32_t* begins[B200TOK_MAX_PEERS];   /* each [world * rows_per_rank]  */
    int32_t* ends[B200TOK_MAX_PEERS];
    int64_t slot_capacity;
    int64_t rows_per_rank;
    /* optional 16-bit wire format (every id < 65 536): ids travel over NVLink as u16 into the peers' staging buffers ids16[p]
     * (each [world * slot_capacity]); after the cross-rank barrier every rank widens its own staging copy into ids[rank] with
     * b200tok_peer_expand_run.  Halves the NVLink bytes; pays one local pass.  ids[p] for p != rank is then unused.     */
    int wire16;
    uint16_t* ids16[B200TOK_MAX_PEERS];
    /* optional NVLS multicast mappings of the same buffers (NULL if unavailable): one multimem.st through the NVSwitch reaches every
     * rank's copy, so each GPU sends its rows once instead of world - 1 times.  Used with the 32-bit wire format.               */
    int32_t* ids_mc;
    int32_t* begins_mc;
    int32_t* ends_mc;
} b200tok_peer_out;
/* n_ids_device (optional, device): this rank's id count.  Fully asynchronous on `cuda_stream`. */
B200TOK_API int b200tok_split_bpe_run_sharded(b200tok_handle split, b200tok_handle bpe, const b200tok_ragged_strings* in,
                                              const b200tok_peer_out* peers, int64_t* n_ids_device, void* cuda_stream);
/* Same for RegexSplit [-> RegexSplit] -> WordpieceTokenizer (rows are compacted, then stored into every rank's slot; with
 * wire16 the caller guarantees every vocabulary id < 65 535). */
B200TOK_API int b200tok_split_wordpiece_run_sharded(b200tok_handle split1, b200tok_handle split2, b200tok_handle wordpiece,
                                                    const b200tok_ragged_strings* in, int32_t unk_token_id, const b200tok_peer_out* peers,
                                                    int64_t* n_ids_device, void* cuda_stream);
/* wire16 only: after the barrier, widen this rank's staging copy (all world * rows_per_rank rows) into peers->ids[rank]. */
B200TOK_API int b200tok_peer_expand_run(int device, const b200tok_peer_out* peers, void* cuda_stream);

/* ---- The same all-gatherv by PULL (replaces the NCCL all-gatherv SURVEY 8e describes, like the calls above) ----------
 * Every rank runs the ordinary b200tok_split_bpe_run / b200tok_split_wordpiece_run on its shard with `out` pointing into
 * peer-mapped buffers (begins / ends / n_ids_device; ids too for the 32-bit wire), optionally packs the ids to 16 bits with
 * b200tok_peer_pack_run, orders ONE cross-rank barrier on the stream, and then gathers with b200tok_peer_pull_run: 16-byte loads
 * from every peer over NVLink, widened straight into the local i32 result.  Layout of the result = the calls above: rank p's rows
 * in slot p (ids[p * slot_capacity ...], begins/ends[p * rows_per_rank + row], offsets shifted by p * slot_capacity).
 * slot_capacity must be a multiple of 8; every source id buffer holds slot_capacity + 8 elements.                          */
typedef struct {
    int world, rank;
    int wire16;                                       /* sources are src_ids16 (every id < 65 536), else src_ids */
    int skip_self_ids;                                /* this rank's ids are already in place in ids[rank * slot_capacity ...] */
    const uint16_t* src_ids16[B200TOK_MAX_PEERS];     /* peer-mapped, each [slot_capacity + 8] */
    const int32_t* src_ids[B200TOK_MAX_PEERS];
    const int32_t* src_begins[B200TOK_MAX_PEERS];     /* peer-mapped, each [rows_per_rank]; offsets relative to that rank's ids */
    const int32_t* src_ends[B200TOK_MAX_PEERS];
    const int64_t* src_total[B200TOK_MAX_PEERS];      /* peer-mapped: that rank's id count */
    int32_t* ids;                                     /* local result [world * slot_capacity] */
    int32_t* begins;                                  /* local result [world * rows_per_rank] */
    int32_t* ends;
    int64_t slot_capacity, rows_per_rank;
} b200tok_peer_pull;
/* ids[0 .. *n_ids_device) (device, i32; readable up to the next multiple of 8) -> ids16 (device, capacity + 8 elements). */
B200TOK_API int b200tok_peer_pack_run(int device, const int32_t* ids, const int64_t* n_ids_device, int64_t capacity, uint16_t* ids16,
                                      void* cuda_stream);
B200TOK_API int b200tok_peer_pull_run(int device, const b200tok_peer_pull* pull, void* cuda_stream);

/* ---- WordpieceTokenizer --------------------------------------------------------------------
 * inputs [5..7] vocab, [8] unk_token_id; attributes suffix_indicator / max_bytes_per_word
 * (src/wordpiece_tokenizer.hpp:41-45). */
typedef struct {
    b200tok_strings vocab;
    const char* suffix_indicator; int64_t suffix_indicator_len;
    int max_bytes_per_word;
    int device;
} b200tok_wordpiece_desc;
B200TOK_API int b200tok_wordpiece_create(const b200tok_wordpiece_desc* desc, b200tok_handle* out);
B200TOK_API int b200tok_wordpiece_run(b200tok_handle h, const b200tok_ragged_strings* in, int32_t unk_token_id,
                                      b200tok_ragged_ids* out, void* cuda_stream);
/* Fused RegexSplit(\s+, remove) -> RegexSplit(punct, isolate) -> WordpieceTokenizer (BERT IR);
 * `split2` may be NULL for a single splitter. */
B200TOK_API int b200tok_split_wordpiece_run(b200tok_handle split1, b200tok_handle split2, b200tok_handle wordpiece,
                                            const b200tok_ragged_strings* in, int32_t unk_token_id,
                                            b200tok_ragged_ids* out, void* cuda_stream);

/* ---- VocabEncoder --------------------------------------------------------------------------
 * inputs [3..5] keys, [6] values (i32 or i64), [7] default (src/vocab_encoder.cpp:56-94). */
typedef struct {
    b200tok_strings keys;
    const void* values;     /* int32_t* or int64_t* */
    int values_are_i64;
    int device;
} b200tok_vocabenc_desc;
B200TOK_API int b200tok_vocabenc_create(const b200tok_vocabenc_desc* desc, b200tok_handle* out);
/* begins/ends: [n]; out: [n] of the value type. */
B200TOK_API int b200tok_vocabenc_run(b200tok_handle h, const int32_t* begins, const int32_t* ends, int64_t n,
                                     const uint8_t* chars, int64_t n_chars, int64_t default_value,
                                     void* out_values, int mem, void* cuda_stream);

/* ---- VocabDecoder (+ fused ByteFallback) -----------------------------------------------------
 * inputs [1..3] vocab; skip tokens come per call (input [4]) or from the attribute
 * (src/vocab_decoder.cpp:36-41). */
typedef struct {
    b200tok_strings vocab;
    int device;
} b200tok_vocabdec_desc;
typedef struct {
    int32_t* ragged_begins;  /* [batch]              */
    int32_t* ragged_ends;    /* [batch]              */
    int32_t* begins;         /* [batch*max(seq,1)]   */
    int32_t* ends;           /* [batch*max(seq,1)]   */
    uint8_t* chars;          /* [chars_capacity]     */
    int64_t chars_capacity;
    int64_t n_chars;         /* out */
    int mem;
} b200tok_decoded;
B200TOK_API int b200tok_vocabdec_create(const b200tok_vocabdec_desc* desc, b200tok_handle* out);
/* ids: i32[batch, seq].  byte_fallback != 0 additionally applies ByteFallback to every decoded
 * token in the same pass (VocabDecoder -> ByteFallback as chained by the detokenizer IR). */
B200TOK_API int b200tok_vocabdec_run(b200tok_handle h, const int32_t* ids, int64_t batch, int64_t seq,
                                     const int32_t* skip_tokens, int64_t n_skip, int byte_fallback,
                                     b200tok_decoded* out, int ids_mem, void* cuda_stream);
/* Upper bound for chars_capacity for a [batch,seq] call (batch*seq*longest vocab entry). */
B200TOK_API int64_t b200tok_vocabdec_max_chars(b200tok_handle h, int64_t batch, int64_t seq);

/* ---- ByteFallback (stand-alone op; stateless) -------------------------------------------------
 * in/out: strings [n]; out_chars worst case n_chars (src/byte_fallback.cpp:25). */
B200TOK_API int b200tok_bytefallback_run(int device, const int32_t* begins, const int32_t* ends, int64_t n,
                                         const uint8_t* chars, int64_t n_chars,
                                         int32_t* out_begins, int32_t* out_ends, uint8_t* out_chars,
                                         int64_t* out_n_chars, int mem, void* cuda_stream);

/* ---- Post-tokenizer tail (stateless): Truncate, CombineSegments, RaggedToDense ----------------------
 * i32 elements (what the converted tokenizer IRs feed these ops: token ids).  `mem` as above.            */
/* Truncate, src/truncate.cpp:37-147.  num_inputs 1: (begins0, ends0); 2: plus (begins1, ends1).  The arrays are
 * edited in place (the reference aliases its outputs to its inputs, :52-56).  side "left" | "right";
 * mode "only_first" | "only_second" | "longest_first" (used with two inputs).                              */
B200TOK_API int b200tok_truncate_run(int device, int num_inputs, int32_t* begins0, int32_t* ends0, int32_t* begins1, int32_t* ends1,
                                     int64_t n, int32_t max_length, const char* side, const char* mode, int mem, void* cuda_stream);
/* One ragged i32 input of CombineSegments: inputs [3j .. 3j+2] of the op. */
typedef struct {
    const int32_t* begins;   /* [n]  */
    const int32_t* ends;     /* [n]  */
    int64_t n;               /* rows; 1 = broadcast to every output row (src/combine_segments.cpp:102-104) */
    const int32_t* elems;    /* [n_elems] */
    int64_t n_elems;
} b200tok_ragged_i32;
/* CombineSegments, src/combine_segments.cpp:36-134.  segment_ids = the op's last input (one id per segment).
 * Outputs: begins/ends [rows] (both output ragged tensors share them, :30-32), elems and ids [capacity];
 * *n_out = produced elements.  At most 16 segments.                                                         */
B200TOK_API int b200tok_combine_segments_run(int device, const b200tok_ragged_i32* segments, int num_segments, const int32_t* segment_ids,
                                             int32_t* out_begins, int32_t* out_ends, int32_t* out_elems, int32_t* out_ids,
                                             int64_t capacity, int64_t* n_out, int mem, void* cuda_stream);
/* RaggedToDense, src/ragged_to_dense.cpp:70-174.  out: i32[n, target_dim]; out_mask: u8[n, target_dim] or NULL.
 * pad_right = attribute or input [5]; pad_max_length = attribute m_pad_max_length.                          */
B200TOK_API int b200tok_ragged_to_dense_run(int device, const int32_t* begins, const int32_t* ends, int64_t n, const int32_t* elems,
                                            int64_t n_elems, int32_t target_dim, int32_t default_value, int pad_right, int pad_max_length,
                                            int32_t* out, uint8_t* out_mask, int mem, void* cuda_stream);
/* The three chained as the converted IRs do for single-sequence inputs — Truncate(max_length, side) ->
 * CombineSegments(prefix constants, tokens, suffix constants) -> RaggedToDense(target_dim, pad_value) — in one pass
 * over the ragged ids; same result as running the three ops back to back.  prefix / suffix: host arrays, <= 8 ids. */
typedef struct {
    int32_t max_length; int truncate_left;
    const int32_t* prefix; int32_t n_prefix;
    const int32_t* suffix; int32_t n_suffix;
    int32_t target_dim; int32_t pad_value; int pad_right;
} b200tok_post_desc;
B200TOK_API int b200tok_post_dense_run(int device, const b200tok_post_desc* desc, const int32_t* begins, const int32_t* ends, int64_t n_rows,
                                       const int32_t* ids, int64_t n_ids, int32_t* out_ids, uint8_t* out_mask, int mem, void* cuda_stream);

/* ---- Byte-level shims and detokenizer tail (stateless) -------------------------------------------
 * BytesToChars, src/bytes_to_chars.cpp:284-339: every byte of every non-skipped element becomes the 1-2 UTF-8 bytes of its
 * GPT-2 printable character; outputs begins/ends [n_elems] and chars (worst case 2 * n_chars); in->skips optional.
 * Rows must cover the elements contiguously and in order (what StringTensorUnpack / RegexSplit produce).               */
B200TOK_API int b200tok_bytes_to_chars_run(int device, const b200tok_ragged_strings* in, int32_t* out_begins, int32_t* out_ends,
                                           uint8_t* out_chars, int64_t chars_capacity, int64_t* n_chars, void* cuda_stream);
/* CharsToBytes, src/chars_to_bytes.cpp:31-68: the inverse map; the ragged dimension is fused (one output string per row):
 * outputs begins/ends [n_rows] and chars (worst case n_chars).                                                         */
B200TOK_API int b200tok_chars_to_bytes_run(int device, const b200tok_ragged_strings* in, int32_t* out_begins, int32_t* out_ends,
                                           uint8_t* out_chars, int64_t chars_capacity, int64_t* n_chars, void* cuda_stream);
/* FuzeRagged, src/fuze.cpp:20-40: out_begins[r] = begins[rb[r]], out_ends[r] = ends[re[r] - 1].                        */
B200TOK_API int b200tok_fuze_ragged_run(int device, const int32_t* ragged_begins, const int32_t* ragged_ends, int64_t n_rows,
                                        const int32_t* begins, const int32_t* ends, int64_t n_elems, int32_t* out_begins,
                                        int32_t* out_ends, int mem, void* cuda_stream);
/* UTF8Validate, src/utf8_validate.cpp:18-137: malformed sequences are replaced by U+FFFD (replace_mode != 0) or dropped.
 * Worst case 3 * n_chars output bytes.  Like the reference, output offsets start at begins[0]; *n_chars = the extent of
 * out_chars in use.                                                                                                    */
B200TOK_API int b200tok_utf8_validate_run(int device, const int32_t* begins, const int32_t* ends, int64_t n, const uint8_t* chars,
                                          int64_t n_chars, int replace_mode, int32_t* out_begins, int32_t* out_ends, uint8_t* out_chars,
                                          int64_t chars_capacity, int64_t* n_chars_out, int mem, void* cuda_stream);

/* ---- normalisers -------------------------------------------------------------------------------------
 * RegexNormalization(global_replace), reference src/regex_normalization.cpp:59-153 + PCRE2Wrapper::substitute
 * src/utils.cpp:315-382.  The GPU path covers the search patterns that match exactly ONE character, i.e. every pattern the
 * BERT normaliser and the prefix / metaspace steps of the converter emit (python/openvino_tokenizers/tokenizer_pipeline.py:
 * 230-278: del_control_chars, replace_whitespace, handle_chinese_chars, strip_accents, add_prefix_whitespace[_to_not_
 * whitespace], prepend, replace_spaces_metaspace and the legacy (^)(.) forms), with a replacement made of literal text and
 * at most one reference to the matched character ($N or \N).  Any other pattern returns B200TOK_E_UNSUPPORTED from
 * _create (nothing runs on the CPU instead).                                                                            */
B200TOK_API int b200tok_regexnorm_create(const char* search_pattern, int64_t search_len, const char* replace_pattern,
                                         int64_t replace_len, int global_replace, int device, b200tok_handle* out);
/* CharsMapNormalization, reference src/charsmap_normalization.cpp:34-69: sentencepiece 0.2.1 normalizer::Normalizer over a
 * precompiled charsmap (u32 trie size | Darts-clone double array | '\0'-separated replacements).  The blob is the op's
 * charsmap input, or what the reference's get_precompiled_charsmap(normalization_form, case_fold) returns for the attribute
 * form.  add_dummy_prefix / remove_extra_whitespaces / escape_whitespaces must be 0 (what NormalizeUnicode / CaseFold
 * set); otherwise B200TOK_E_UNSUPPORTED.                                                                                */
B200TOK_API int b200tok_charsmap_create(const uint8_t* precompiled_charsmap, int64_t charsmap_len, int add_dummy_prefix,
                                        int remove_extra_whitespaces, int escape_whitespaces, int device, b200tok_handle* out);
/* evaluate_normalization_helper, reference src/utils.cpp:178-234, for either handle: strings (begins, ends, chars) ->
 * normalised strings packed back to back from offset 0; elements with skips[i] != 0 are copied unchanged (skips may be
 * NULL).  Worst case output: RegexNormalization n_chars * (1 + replacement length), CharsMapNormalization bounded by the
 * longest replacement per input byte; a too small chars_capacity returns B200TOK_E_CAPACITY with *n_chars_out = the size
 * needed.  Free the handle with b200tok_destroy.                                                                        */
B200TOK_API int b200tok_normalize_run(b200tok_handle h, const int32_t* begins, const int32_t* ends, int64_t n, const uint8_t* chars,
                                      int64_t n_chars, const uint8_t* skips, int32_t* out_begins, int32_t* out_ends, uint8_t* out_chars,
                                      int64_t chars_capacity, int64_t* n_chars_out, int mem, void* cuda_stream);
/* The same for a chain of normalisers applied one after the other (handles[0] first) — what the converter emits for one
 * HF normaliser, e.g. BertNormalizer = del_control_chars, replace_whitespace, handle_chinese_chars, NFD, strip_accents,
 * case fold (python/openvino_tokenizers/hf_parser.py:84-102 of the reference).  Intermediate strings stay on the device;
 * skips applies to every op, as in the reference graph where each op passes the tensor on (src/utils.cpp:189-191).       */
B200TOK_API int b200tok_normalize_chain_run(const b200tok_handle* handles, int n_ops, const int32_t* begins, const int32_t* ends, int64_t n,
                                            const uint8_t* chars, int64_t n_chars, const uint8_t* skips, int32_t* out_begins,
                                            int32_t* out_ends, uint8_t* out_chars, int64_t chars_capacity, int64_t* n_chars_out, int mem,
                                            void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* B200TOK_H_ */
