/* minorg_b200 - C ABI of the B200-native gRNA enumeration + off-target screen.
 *
 * The reference (rlrq/MINORg, pure Python) has no FFI of its own: the seam is Python-level and the
 * arithmetic is delegated to `regex` (PAM scan) and to a `blastn` subprocess (screen, masks).  Each entry
 * point below names the reference call it stands in for (paths under /root/reference/minorg/).
 * INTEGRATION.md shows the ctypes binding a MINORg maintainer would add at each of those call sites.
 *
 * Conventions: every function returns 0 on success or a negative mg_status; the message of the last
 * failure on the calling thread is mg_last_error().  Pointers named d_* are DEVICE pointers, all others
 * are host pointers.  `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 * Handles are owned by the library (mg_*_free).  Not re-entrant per handle.  Do not fork() after the
 * first call.  Coordinates are 0-based, end-exclusive, per contig, forward strand (Biopython HSP
 * convention, filter_grna.py:52-56).
 */
#ifndef MINORG_B200_H
#define MINORG_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    MG_OK = 0,
    MG_ERR_CUDA = -1,      /* a CUDA runtime call or kernel failed */
    MG_ERR_ARG = -2,       /* invalid argument */
    MG_ERR_NOMEM = -3,     /* host or device allocation failed */
    MG_ERR_OVERFLOW = -4,  /* caller-provided output capacity too small (needed size is reported) */
    MG_ERR_UNSUPPORTED = -5
} mg_status;

#define MG_MAX_GRNA_LEN 32   /* gRNA length limit of the device kernels */
#define MG_MAX_PAM_LEN 16    /* longest fixed-length PAM alternative */
#define MG_MAX_PAM_ALTS 32   /* alternatives a quantified PAM side may expand to */
#define MG_MAX_PATTERN_UNITS 32
#define MG_MAX_PATTERN_OPS 96
#define MG_MAX_GAPS 3
#define MG_MAX_PIECES 8

const char* mg_last_error(void);
int mg_version(void);
int mg_device_count(void);
/* Select the device for the calling thread and report its name / SM count (either may be NULL). */
int mg_init(int device, char* name_out, int name_cap, int* sm_count_out);

/* ------------------------------------------------------------------------------------------------
 * Genome store: 2-bit packed bases + 1-bit "invalid" plane resident in HBM.
 * Stands in for the subject side of the reference's BLAST call and for IndexedFasta slicing
 * (index.py:150-211; MINORg.py:2134-2136, 2066-2083).  Contigs are laid out in one global coordinate
 * space separated by >= 64 invalid bases, so no window spans two contigs and positions beyond a contig
 * end read as "mismatch with everything".  Any byte other than ACGTacgt is invalid.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mg_genome mg_genome;

/* Native FASTA reader (host threads; plain files): what the reference does with Bio.SeqIO / pyfaidx before it can
 * hand a subject to blastn (fasta.py:21-43, index.py:150-211).  A header line starts with '>'; the record id is the
 * first whitespace-delimited token of the header; the sequence is every byte up to the next header line except
 * '\n', '\r', ' ' and '\t' (case preserved); bytes before the first header are ignored.  threads <= 0: all cores. */
typedef struct mg_fasta mg_fasta;
int mg_fasta_read(const char* path, int threads, mg_fasta** out);
int mg_fasta_n_records(const mg_fasta* f);
const char* mg_fasta_name(const mg_fasta* f, int i);
const int64_t* mg_fasta_offsets(const mg_fasta* f);      /* n_records + 1 */
const char* mg_fasta_sequence(const mg_fasta* f);        /* offsets[n_records] bytes, records concatenated */
void mg_fasta_free(mg_fasta* f);

/* seq: concatenated ASCII contig sequences (host); contig i = seq[offsets[i], offsets[i+1]).  */
int mg_genome_from_ascii(const char* seq, const int64_t* offsets, int n_contigs, void* stream, mg_genome** out);
/* mg_fasta_read + mg_genome_from_ascii.  fasta_out (may be NULL) receives the parsed file (names, offsets, host
 * sequence) - the caller frees it with mg_fasta_free. */
int mg_genome_from_fasta(const char* path, int threads, void* stream, mg_genome** out, mg_fasta** fasta_out);
/* Same, with the ASCII bytes already on the device (synthetic-genome benchmarks). */
int mg_genome_from_device_ascii(const char* d_seq, const int64_t* offsets, int n_contigs, void* stream,
                                mg_genome** out);
/* Chunk of a subject for one GPU of several (BASELINE.json north_star: "genomes are sharded by chunk (with gRNA-length
 * halo) across the 8 GPUs").  The coordinate space is the WHOLE subject's - seq/offsets/n_contigs
 * describe all of it, and contigs, masks and window ranges keep their global meaning - but only the bases of global
 * positions [chunk_begin - halo, chunk_end + halo) are copied to the device and packed.  The handle owns the windows
 * whose first base lies in [chunk_begin, chunk_end): mg_screen and mg_mask_find restrict themselves to them, so the
 * chunks of a partition of [0, mg_layout_global_length()) evaluate every cell (and report every mask occurrence)
 * exactly once.  halo >= 256 and >= the longest mask sequence + 64 when mg_mask_find is used.  The reference has no
 * counterpart: it hands whole FASTA files to blastn processes (MINORg.py:2182-2195). */
int64_t mg_layout_global_length(const int64_t* offsets, int n_contigs);   /* global length of the layout, no device work */
int mg_genome_from_ascii_chunk(const char* seq, const int64_t* offsets, int n_contigs, int64_t chunk_begin,
                               int64_t chunk_end, int64_t halo, void* stream, mg_genome** out);
int mg_genome_from_device_ascii_chunk(const char* d_seq, const int64_t* offsets, int n_contigs, int64_t chunk_begin,
                                      int64_t chunk_end, int64_t halo, void* stream, mg_genome** out);
/* owned window range and stored base range (global coordinates); any pointer may be NULL */
int mg_genome_owned_range(const mg_genome* g, int64_t* begin, int64_t* end, int64_t* stored_begin, int64_t* stored_end);
int mg_genome_n_contigs(const mg_genome* g);
int64_t mg_genome_contig_length(const mg_genome* g, int contig);
int64_t mg_genome_total_bases(const mg_genome* g);    /* sum of contig lengths (no padding) */
int64_t mg_genome_device_bytes(const mg_genome* g);   /* HBM footprint of the packed planes */
/* Decode bases [start,end) of a contig back to ASCII ('N' for invalid); test/debug aid. */
int mg_genome_decode(const mg_genome* g, int contig, int64_t start, int64_t end, char* out);
void mg_genome_free(mg_genome* g);

/* Masked (on-target) intervals of this genome: replaces MINORg.masked[filename] + _is_offtarget_pos
 * (MINORg.py:1953-1964, 1994-2000).  A hit is on-target iff its span lies inside ONE interval. */
int mg_genome_set_masks(mg_genome* g, const int32_t* contig, const int64_t* start, const int64_t* end, int n,
                        void* stream);

/* K3 - exact-occurrence mask finder: replaces mask_identical/blast_mask (filter_grna.py:67-119):
 * every interval of the genome equal to a query sequence end-to-end, on either strand (ACGT only,
 * case-insensitive).  seqs = concatenated ASCII, sequence i = seqs[offsets[i], offsets[i+1]).
 * Outputs (host arrays of capacity cap): sequence index, contig, start, end, strand (+1/-1).
 * *n_found receives the number found even when it exceeds cap (then MG_ERR_OVERFLOW). */
int mg_mask_find(const mg_genome* g, const char* seqs, const int64_t* offsets, int n_seqs, int64_t cap,
                 int32_t* out_seq, int32_t* out_contig, int64_t* out_start, int64_t* out_end, int8_t* out_strand,
                 int64_t* n_found, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PAM program: a PAM side ('five' = pattern before the gRNA, 'three' = after) expanded by the host to a
 * finite union of fixed-length alternatives; position k of an alternative allows the ASCII bytes whose
 * bit is set in allow[alt][k][byte>>5] (enumeration; case already folded in by the host) and the bases
 * whose bit is set in allow4[alt][k] (bit 0..3 = A,C,G,T; PAM proximity check on the 2-bit genome).
 * Mirrors what PAM.regex()/five_prime()/three_prime() hand to `regex` (pam.py:103-125, 150-167).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t n_alts;                                   /* 0 = this side is empty */
    int32_t len[MG_MAX_PAM_ALTS];
    uint32_t allow[MG_MAX_PAM_ALTS][MG_MAX_PAM_LEN][8];
    uint8_t allow4[MG_MAX_PAM_ALTS][MG_MAX_PAM_LEN];
} mg_pam_side;

typedef struct {
    mg_pam_side five;
    mg_pam_side three;
    int32_t five_maxlen;    /* PAM.five_prime_maxlen(), including its documented over-count (pam.py:128-147) */
    int32_t three_maxlen;
} mg_pam;

/* K2 - PAM-site enumeration: replaces the regex.finditer(overlapped=True) scan of both strands in
 * find_multi_gRNA (generate_grna.py:38-41).  seq = one target record (ASCII, case preserved).  Outputs
 * the ascending match starts on the '+' strand and on the '-' strand (positions in the reverse-complemented
 * record, exactly like the reference's loop); capacities cap each; counts are reported even on overflow. */
int mg_enumerate_sites(const char* seq, int64_t len, const mg_pam* pam, int grna_len, int64_t cap,
                       int64_t* out_plus, int64_t* n_plus, int64_t* out_minus, int64_t* n_minus, void* stream);

/* ------------------------------------------------------------------------------------------------
 * gRNA set resident in HBM: N sequences of one length L (ASCII, any case; non-ACGT = mismatch with
 * everything).  Holds both orientations (query i = gRNA i, query N+i = its reverse complement) and the
 * bit-sliced mismatch tables of the scan kernel.  Replaces the gRNA FASTA handed to blastn
 * (MINORg.py:2134-2135).
 * ---------------------------------------------------------------------------------------------- */
typedef struct mg_queries mg_queries;
int mg_queries_create(const char* grna, int n, int len, void* stream, mg_queries** out);
int mg_queries_count(const mg_queries* q);
int mg_queries_length(const mg_queries* q);
void mg_queries_free(mg_queries* q);

/* Off-target criteria = the reference attributes that decide a hit (MINORg.py:686-707, 2018-2102).
 * max_mismatch = max(1, ot_mismatch) - 1, max_gap = max(1, ot_gap) - 1 (MINORg.py:2033-2034).
 * Pattern mode (n_units > 0) replaces OffTargetExpression (ot_regex.py:124-356): unit u is true iff
 * count(chars in slice) [+ unaligned] <= n; `ops` is the expression in postfix order:
 * op >= 0 pushes unit op; MG_OP_AND / MG_OP_OR combine the two topmost values. */
#define MG_OP_AND (-1)
#define MG_OP_OR (-2)
#define MG_CH_MISMATCH 1
#define MG_CH_INS 2
#define MG_CH_DEL 4
#define MG_SLICE_NONE INT32_MIN
typedef struct {
    int32_t n;          /* threshold */
    int32_t chars;      /* MG_CH_* mask of characters counted */
    int32_t start, end; /* Python slice bounds from ot_range (MG_SLICE_NONE = None) */
    int32_t rvs;        /* 1: slice the tuple whose deletions join the PREVIOUS position (blast.py:102-117) */
    int32_t add_unaligned_m, add_unaligned_g; /* unaligned_as_mismatch / unaligned_as_gap apply to this unit */
} mg_pattern_unit;

typedef struct {
    int32_t max_mismatch;
    int32_t max_gap;
    int32_t pamless;          /* 1 = skip the PAM proximity check (reference default) */
    int32_t n_units;          /* 0 = no pattern (MINORg._is_offtarget_aln_nopattern) */
    int32_t n_ops;
    mg_pattern_unit units[MG_MAX_PATTERN_UNITS];
    int32_t ops[MG_MAX_PATTERN_OPS];
    int32_t prefilter_edits;  /* host-derived bound: pattern true => full-length edit cost <= this; <0 = unbounded */
    /* Optional exact-piece prefilter for patterns (pigeonhole): the host proves that every alignment satisfying
     * the pattern matches the subject EXACTLY, without gaps, on at least one of these gRNA position ranges
     * (in gRNA orientation).  n_pieces = 0: not used. */
    int32_t n_pieces;
    int32_t piece_start[MG_MAX_PIECES];
    int32_t piece_len[MG_MAX_PIECES];
} mg_criteria;

/* K4/K5 - exhaustive off-target screen of every resident gRNA against every window of the genome, both
 * strands: replaces MINORg._offtarget_hits (blastn -task blastn-short + is_offtarget_hit per HSP,
 * MINORg.py:2123-2144).  d_counts: DEVICE array of 2*N uint32 that is ADDED to (query i / N+i = plus /
 * minus orientation of gRNA i): the number of problematic, unmasked alignments anchored in this genome.
 * gRNA i fails the background check iff d_counts[i] + d_counts[N+i] > 0.  No early exit: every cell is
 * evaluated, and the counts are exact in every mode: each problematic anchor (strand, virtual alignment end) is counted once,
 * whatever prefilter found it and however the window range is cut into shards or slices.  window range [win_begin, win_end) in global coordinates selects a shard (pass 0, -1 for all).
 * pam may be NULL when crit->pamless.
 * Algorithm by criterion (same counts whichever runs): no gap -> bit-sliced pair-word scan; gaps or a pattern -> a prefilter
 * (bit-sliced edit distance, or exact pieces) + the exact verifier; no pattern, <= 1 gap, <= 4 mismatches (<= 5 without gap),
 * PAM-less, 20-nt gRNA and a large set (>= 25 000 gRNA with a gap, >= 100 000 without; environment MG_SEEDJOIN_MIN, and
 * MG_SEEDJOIN=off|on forces either) -> the seed-and-join screen, which builds a hash index of the gRNA set on first use and
 * keeps it in the mg_queries handle (~170 B per gRNA). */
int mg_screen(const mg_genome* g, const mg_queries* q, const mg_criteria* crit, const mg_pam* pam,
              int64_t win_begin, int64_t win_end, uint32_t* d_counts, void* stream);

/* mg_screen on resident handles with the 2*N counts returned in HOST memory (device scratch inside the library):
 * what the Python mirror of MINORg._offtarget_hits calls, so that a single-GPU run needs no device-array package. */
int mg_screen_to_host(const mg_genome* g, const mg_queries* q, const mg_criteria* crit, const mg_pam* pam,
                      int64_t win_begin, int64_t win_end, uint32_t* counts_out, void* stream);

/* Host-buffer convenience around the calls above (the end-to-end path timed by bench.py "e2e"):
 * ASCII genome + ASCII gRNAs in host memory -> H2D, pack, screen, D2H.  counts_out: N uint32 (both
 * orientations summed).  The subject is processed in batches of whole contigs; the next batch is copied and
 * packed on an internal stream while the current one is screened on `stream` (pinned `seq` makes that copy
 * asynchronous).  Returns after everything, including the D2H of the counts, has completed.  The query tables of the
 * previous call on this thread are kept and reused when the gRNA bytes are the same (a pipeline screens one gRNA set against
 * subject after subject, MINORg.py:2198-2207), as are the internal copy stream and the pooled device scratch. */
int mg_screen_host(const char* seq, const int64_t* offsets, int n_contigs,
                   const int32_t* mask_contig, const int64_t* mask_start, const int64_t* mask_end, int n_masks,
                   const char* grna, int n, int len, const mg_criteria* crit, const mg_pam* pam,
                   uint32_t* counts_out, void* stream);

/* Layout of the global coordinate space (for sharding by chunk with a halo): global start of a contig. */
int64_t mg_genome_contig_global_start(const mg_genome* g, int contig);
int64_t mg_genome_global_length(const mg_genome* g);

/* Work actually executed by the most recent bit-sliced scan launch on this device (synchronises):
 * out[0] = (warp, 32-query group) steps, out[1] = how many of them went past the early-out test, i.e. also
 * loaded and summed the last quarter of the gRNA positions.  Used by bench.py for the roofline accounting. */
int mg_last_screen_stats(uint64_t* out /* [2] */);

/* Work of the most recent seed-and-join screen on this device (the one-gap criterion on a large gRNA set, see
 * mg_screen; synchronises): out[0] = candidates, i.e. index entries compared with the text, out[1] = candidates that
 * passed the first (necessary) test and went through the exact evaluation.  Zero if that path did not run. */
int mg_last_seedjoin_stats(uint64_t* out /* [2] */);

/* Integer-pipe microbenchmarks used for the roofline denominators of the scan kernel (bench.py):
 * which: 0 = LOP3 (ALU pipe) lane-ops, 1 = shared-memory 32-bit loads (LSU), 2 / 3 = 64- / 128-bit shared-memory
 * loads with the scan's broadcast pattern (4 adjacent entries per warp), 4 = indexed constant-bank loads, 5 / 6 = warp
 * shuffles alone / interleaved 1:1 with shared loads, 7 = IMAD (FMA pipe), 8 = IMAD and LOP3 interleaved 1:1 (both integer
 * pipes together: the issue-slot ceiling of the gapped rows kernel).  Returns G lane-instructions/s in *gops. */
int mg_microbench(int which, double* gops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MINORG_B200_H */
