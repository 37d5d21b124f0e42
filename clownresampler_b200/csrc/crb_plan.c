/*
 * crb_plan.c -- host-side construction of the per-phase tap table ("plan") the CUDA kernels use.
 *
 * The reference (H = /root/reference/clownresampler.h) recomputes, for every output frame, the
 * tap window bounds (H:993-996), the table index of the first tap (H:1001), walks the 6144-entry
 * Lanczos table with stride kernel_step_size (H:1008-1016) and divides 0x80000000 by the tap sum
 * (H:1025).  All of that depends only on the 16-bit position fraction.  This file re-lays the
 * table out by phase once per configuration:
 *
 *   e      = ceil(q / 65536) * 65536 - q,   q = position(16.16) + radius_delta        (0..65535)
 *   ks(e)  = (step * (e + delta)) >> 16      == kernel_start of H:1001
 *   row(e) = ks(e) - ks(0) + #{break b : e >= b}
 *   row    = [ |k| of column 0 .. n_cols-1 ][ 0x80000000 / sum(k) ][pad]
 *
 * Columns are the taps of the widest window, in input order, with the weight SIGN made static
 * per column: a tap index whose weight is positive in some phases and negative in others is
 * split into a (+) and a (-) column, taps that are zero in every phase are dropped.  Runs of
 * equal-sign columns let the kernel accumulate |k|-weighted truncated products in two chains
 * (positive, negative) whose truncation bias depends on the sample sign only -- see
 * crb_device.cu.  Every structural assumption is verified here by brute force over all 65536
 * fractions; a configuration that violates one is rejected, never approximated.
 */
#include "crb_internal.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread char g_error[512];

/* CRB200_FORCE_DIRECT=1 makes every plan use the direct global-memory kernel (test hook). */
static int force_direct(void)
{
	const char *v = getenv("CRB200_FORCE_DIRECT");
	return v && v[0] == '1';
}

/* CRB200_NO_SMALL=1 keeps slightly stretched kernels on the general kernel (test / A-B hook). */
static int no_small(void)
{
	const char *v = getenv("CRB200_NO_SMALL");
	return v && v[0] == '1';
}

/* CRB200_NO_CHAINS=1 keeps the general kernel on its IMAD.HI form (test / A-B hook). */
static int no_chains(void)
{
	const char *v = getenv("CRB200_NO_CHAINS");
	return v && v[0] == '1';
}

typedef struct chain_col {
	uint32_t col;                /* column before regrouping */
	int64_t maxabs;              /* largest |k| over the phase rows */
} chain_col;

static int chain_col_cmp(const void *a, const void *b)
{
	const chain_col *x = (const chain_col *)a, *y = (const chain_col *)b;
	if (x->maxabs != y->maxabs) return x->maxabs > y->maxabs ? -1 : 1;
	return x->col < y->col ? -1 : x->col > y->col;
}

void crb_set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_error, sizeof g_error, fmt, ap);
	va_end(ap);
}

const char *ClownResamplerB200_GetLastError(void)
{
	return g_error;
}

uint64_t crb_hash_table(const long *table)
{
	/* FNV-1a over the 6144 values (as 64-bit words); identifies a caller's table contents */
	uint64_t h = 1469598103934665603ull;
	size_t i;
	for (i = 0; i < CRB_TABLE_SIZE; ++i) {
		h ^= (uint64_t)table[i];
		h *= 1099511628211ull;
	}
	return h;
}

/* ---- shared-memory bank model used to pick the thread -> frame stride and the column rotation ---- */

/* wavefronts of one warp-wide load: lane l reads `width` bytes at byte address addr[l]; 8- and 16-byte loads go
   in half- and quarter-warp passes; lanes of a pass that touch different 32-bit words of one bank serialise */
static uint32_t load_wavefronts(const uint64_t addr[32], uint32_t width)
{
	const uint32_t lanes_per_pass = width <= 4 ? 32 : width == 8 ? 16 : 8;
	const uint32_t words = width <= 4 ? 1 : width / 4;
	uint32_t total = 0, pass;
	for (pass = 0; pass < 32 / lanes_per_pass; ++pass) {
		uint64_t word_in_bank[32][8];
		uint32_t count[32], l, b, worst = 0;
		memset(count, 0, sizeof count);
		for (l = 0; l < lanes_per_pass; ++l) {
			const uint64_t w0 = addr[pass * lanes_per_pass + l] / 4;
			uint32_t k;
			for (k = 0; k < words; ++k) {
				const uint64_t w = w0 + k;
				uint32_t i, seen = 0;
				b = (uint32_t)(w & 31);
				for (i = 0; i < count[b]; ++i) if (word_in_bank[b][i] == w) seen = 1;
				if (!seen && count[b] < 8) word_in_bank[b][count[b]++] = w;
			}
		}
		for (b = 0; b < 32; ++b) if (count[b] > worst) worst = count[b];
		total += worst;
	}
	return total;
}

typedef struct bank_model {
	uint64_t increment;
	uint32_t channels, threads, step, delta, ks0, n_breaks;
	const uint32_t *breaks;
	/* the group the model watches (the widest): its columns' frame offsets, and where it sits in a row */
	const uint32_t *col_off;     /* bytes, per column of the group */
	uint32_t count, first, row_words;
} bank_model;

/* average wavefronts of one loop iteration (two sample loads, one weight pair, one offset pair) of the watched group
   when thread t takes frame (t * lane_stride) mod threads and lane l starts at pair ((rot * l) >> rot_shift) & rot_mask */
static double iteration_wavefronts(const bank_model *m, uint32_t lane_stride, uint32_t rot, uint32_t rot_shift, uint32_t rot_mask)
{
	const uint32_t frame_bytes = 2 * m->channels;
	const uint32_t width = m->channels == 2 ? 4 : m->channels == 4 ? 8 : m->channels == 8 ? 16 : 2;
	const uint32_t pairs = m->count / 2;
	uint64_t total = 0;
	uint32_t trials = 0, phase, warp, p;
	for (phase = 0; phase < 65536; phase += 16411)
		for (warp = 0; warp < 4; ++warp)
			for (p = 0; p < pairs; p += (pairs + 2) / 3, ++trials) {
				uint64_t a0[32], a1[32], aw[32], ao[32];
				uint32_t l;
				for (l = 0; l < 32; ++l) {
					const uint32_t t = warp * 32 + l;
					const uint32_t f = (t * lane_stride) & (m->threads - 1);
					const uint64_t q = phase + (uint64_t)f * m->increment;
					const uint64_t ws = (q + 65535) >> 16;
					const uint32_t e = (uint32_t)((ws << 16) - q);
					uint32_t row = (uint32_t)(((uint64_t)m->step * (e + m->delta)) >> 16) - m->ks0, b, pp;
					for (b = 0; b < m->n_breaks; ++b) row += (e >= m->breaks[b]);
					const uint32_t sl = ((rot * l) >> rot_shift) & rot_mask;
					pp = p + sl;
					if (pp >= pairs) pp -= pairs;        /* the copy behind the group holds the same columns */
					a0[l] = ws * frame_bytes + m->col_off[2 * pp];
					a1[l] = ws * frame_bytes + m->col_off[2 * pp + 1];
					aw[l] = ((uint64_t)row * m->row_words + m->first + 2 * (p + sl)) * 4;
					ao[l] = ((uint64_t)m->first + 2 * (p + sl)) * 4;
				}
				total += load_wavefronts(a0, width) + load_wavefronts(a1, width) + load_wavefronts(aw, 8) + load_wavefronts(ao, 8);
			}
	return (double)total / trials;
}

/* average wavefronts of one sample load of the slightly stretched kernel (consecutive frames on consecutive lanes;
   16-bit loads per channel, or packed 32-bit loads per channel pair for even channel counts with up to 8 taps) */
static double small_kernel_load_wavefronts(uint64_t increment, uint32_t channels, uint32_t taps)
{
	const uint32_t width = (channels % 2 == 0 && taps <= 8) ? 4 : 2;
	uint64_t total = 0;
	uint32_t trials = 0, phase, warp, l;
	for (phase = 0; phase < 65536; phase += 16411)
		for (warp = 0; warp < 4; ++warp, ++trials) {
			uint64_t a[32];
			for (l = 0; l < 32; ++l) {
				const uint64_t q = phase + (uint64_t)(warp * 32 + l) * increment;
				a[l] = ((q + 65535) >> 16) * 2 * channels;
			}
			total += load_wavefronts(a, width);
		}
	return (double)total / trials;
}

typedef struct phase_key {
	uint32_t ks, ntaps;
} phase_key;

int crb_plan_build_host(struct ClownResamplerB200_Plan *plan, const long *table,
	uint64_t radius_fx, uint64_t radius_int, uint64_t delta, uint64_t step,
	uint64_t increment, unsigned channels, uint32_t smem_budget_bytes)
{
	crb_geometry *g = &plan->geo;
	phase_key *by_e = NULL, *row_key = NULL;
	uint8_t *tap_pos = NULL, *tap_neg = NULL;
	int32_t *col_of_tap_pos = NULL, *col_of_tap_neg = NULL;
	uint32_t *row_of_e_start = NULL;
	uint32_t frac, e, i, r, n_rows, taps_max = 0, n_cols, n_runs;
	uint64_t taps_total = 0;
	int norm_t16_ok = 1;
	int rc = -3; /* CRB200_E_CONFIG */

	memset(g, 0, sizeof *g);
	if (channels == 0 || channels > CRB_MAX_CHANNELS) {
		crb_set_error("channels must be 1..%d (got %u); the reference's accumulator array has %d slots (H:1071)", CRB_MAX_CHANNELS, channels, CRB_MAX_CHANNELS);
		return rc;
	}
	if (step == 0) {
		crb_set_error("kernel_step_size is 0: every tap would read table[0] == 0 and the reference divides by a zero tap sum (H:981, H:1025)");
		return rc;
	}
	if (increment == 0 || increment > 0xFFFFFFFFull || radius_int == 0 || delta >= CRB_FX_ONE || step > 1024
	    || radius_fx + delta != radius_int * CRB_FX_ONE || radius_fx >= ((uint64_t)1 << 32)) {
		crb_set_error("inconsistent resampler state (increment=%llu radius_fx=%llu radius_int=%llu delta=%llu step=%llu); was it initialised with ClownResampler_LowLevel_Init?",
			(unsigned long long)increment, (unsigned long long)radius_fx, (unsigned long long)radius_int, (unsigned long long)delta, (unsigned long long)step);
		return rc;
	}

	plan->host_table = (int32_t *)malloc(CRB_TABLE_SIZE * sizeof(int32_t));
	by_e = (phase_key *)malloc(65536 * sizeof *by_e);
	if (!plan->host_table || !by_e) { crb_set_error("out of host memory"); rc = -5; goto fail; }
	for (i = 0; i < CRB_TABLE_SIZE; ++i) {
		if (table[i] < -0x7FFFFFFFl || table[i] > 0x7FFFFFFFl) {
			crb_set_error("kernel table entry %u = %ld does not fit the 32-bit device table", i, table[i]);
			goto fail;
		}
		plan->host_table[i] = (int32_t)table[i];
	}

	/* 1. the reference's window geometry for every fraction (H:993-1001) */
	for (frac = 0; frac < 65536; ++frac) {
		const uint64_t min_rel = (frac + delta + (CRB_FX_ONE - 1)) / CRB_FX_ONE;
		const uint64_t max_rel = (frac + radius_fx) / CRB_FX_ONE;
		const uint64_t ntaps = radius_int + max_rel - min_rel;
		const uint64_t u = min_rel * CRB_FX_ONE - frac;
		const uint64_t ks = step * u / CRB_FX_ONE;
		if (min_rel > radius_int || max_rel > radius_int || ntaps == 0) {     /* H:1003-1004 */
			crb_set_error("window bounds exceed the kernel radius at fraction %u (the reference asserts here)", frac);
			goto fail;
		}
		if (ks + (ntaps - 1) * step >= CRB_TABLE_SIZE) {                        /* H:1012 */
			crb_set_error("kernel table index %llu out of range at fraction %u (the reference asserts here)",
				(unsigned long long)(ks + (ntaps - 1) * step), frac);
			goto fail;
		}
		e = (uint32_t)(u - delta);      /* == ceil(q/65536)*65536 - q with q = frac + delta */
		if (e > 65535) { crb_set_error("internal: phase coordinate out of range"); goto fail; }
		by_e[e].ks = (uint32_t)ks;
		by_e[e].ntaps = (uint32_t)ntaps;
		if (ntaps > taps_max) taps_max = (uint32_t)ntaps;
		taps_total += ntaps;
	}
	plan->mean_taps = (double)taps_total / 65536.0;

	/* 2. rows: maximal runs of e with the same (ks, ntaps); ks must advance by single steps */
	row_key = (phase_key *)malloc(65536 * sizeof *row_key);
	row_of_e_start = (uint32_t *)malloc(65536 * sizeof *row_of_e_start);
	if (!row_key || !row_of_e_start) { crb_set_error("out of host memory"); rc = -5; goto fail; }
	g->ks0 = by_e[0].ks;
	n_rows = 1;
	row_key[0] = by_e[0];
	row_of_e_start[0] = 0;
	for (e = 1; e < 65536; ++e) {
		if (by_e[e].ks != by_e[e - 1].ks) {
			if (by_e[e].ks != by_e[e - 1].ks + 1) { crb_set_error("internal: kernel_start is not a unit-step function of the phase"); goto fail; }
		} else if (by_e[e].ntaps != by_e[e - 1].ntaps) {
			if (g->n_breaks == CRB_MAX_BREAKS) { crb_set_error("internal: too many tap-count breakpoints"); goto fail; }
			g->breaks[g->n_breaks++] = e;
		} else {
			continue;
		}
		row_key[n_rows] = by_e[e];
		row_of_e_start[n_rows] = e;
		++n_rows;
	}

	/* 3. sign and magnitude of every tap index over all rows */
	tap_pos = (uint8_t *)calloc(taps_max, 1);
	tap_neg = (uint8_t *)calloc(taps_max, 1);
	col_of_tap_pos = (int32_t *)malloc(taps_max * sizeof(int32_t));
	col_of_tap_neg = (int32_t *)malloc(taps_max * sizeof(int32_t));
	if (!tap_pos || !tap_neg || !col_of_tap_pos || !col_of_tap_neg) { crb_set_error("out of host memory"); rc = -5; goto fail; }
	for (r = 0; r < n_rows; ++r)
		for (i = 0; i < row_key[r].ntaps; ++i) {
			const int32_t k = plan->host_table[row_key[r].ks + i * step];
			/* 1 = present and every |k| < 32768 ("small"), 2 = present with some |k| >= 32768 ("big") */
			if (k > 0 && tap_pos[i] < 2) tap_pos[i] = k >= 32768 ? 2 : 1;
			if (k < 0 && tap_neg[i] < 2) tap_neg[i] = -(int64_t)k >= 32768 ? 2 : 1;
		}

	/* 4. columns and runs of equal (sign class, form), in input order.  A tap whose weight is positive in some phase rows
	      and negative in others (a zero crossing of the stretched kernel passes over it) becomes one signed column. */
	n_cols = 0;
	n_runs = 0;
	for (i = 0; i < taps_max; ++i) {
		int neg, big;
		crb_run *last = n_runs ? &g->runs[n_runs - 1] : NULL;
		col_of_tap_pos[i] = col_of_tap_neg[i] = -1;
		if (!tap_pos[i] && !tap_neg[i]) continue;
		neg = tap_pos[i] && tap_neg[i] ? 2 : tap_neg[i] ? 1 : 0;
		big = tap_pos[i] == 2 || tap_neg[i] == 2;
		if (last && last->negative == neg && last->big == big && (uint32_t)(last->off + last->len) == i && (uint32_t)(last->col + last->len) == n_cols) {
			++last->len;
		} else {
			if (n_runs == CRB_MAX_RUNS) { crb_set_error("internal: too many sign runs in the kernel"); goto fail; }
			g->runs[n_runs].col = (int32_t)n_cols;
			g->runs[n_runs].len = 1;
			g->runs[n_runs].off = (int32_t)i;
			g->runs[n_runs].negative = (int16_t)neg;
			g->runs[n_runs].big = (int16_t)big;
			++n_runs;
		}
		if (tap_pos[i]) col_of_tap_pos[i] = (int32_t)n_cols;
		if (tap_neg[i]) col_of_tap_neg[i] = (int32_t)n_cols;
		++n_cols;
	}
	if (n_cols == 0) { crb_set_error("kernel has no non-zero taps"); goto fail; }
	g->n_cols = n_cols;
	g->n_runs = n_runs;
	g->row_words = (n_cols + 1) | 1u;   /* odd stride: lanes reading column c of different rows spread over the banks */
	g->taps_max = taps_max;

	/* 5. row contents + reciprocal (H:1025) + range proofs for the 32-bit device arithmetic.
	      small columns hold |k| << 16, big columns |k|; the last word holds the reciprocal */
	plan->host_rows = (int32_t *)calloc((size_t)n_rows * g->row_words, sizeof(int32_t));
	if (!plan->host_rows) { crb_set_error("out of host memory"); rc = -5; goto fail; }
	g->norm_mode = 3;
	for (r = 0; r < n_rows; ++r) {
		int32_t *row = plan->host_rows + (size_t)r * g->row_words;
		int64_t sum = 0, sum_pos = 0, sum_neg = 0, recip;
		for (i = 0; i < row_key[r].ntaps; ++i) {
			const int64_t k = plan->host_table[row_key[r].ks + i * step];
			sum += k;
			const int is_signed = tap_pos[i] && tap_neg[i];
			const int big = is_signed ? (tap_pos[i] == 2 || tap_neg[i] == 2) : k > 0 ? tap_pos[i] == 2 : tap_neg[i] == 2;
			if (k > 0) { row[col_of_tap_pos[i]] = (int32_t)(big ? k : k * 65536); sum_pos += k; }
			if (k < 0) { row[col_of_tap_neg[i]] = (int32_t)(is_signed ? (big ? k : k * 65536) : (big ? -k : -k * 65536)); sum_neg -= k; }
		}
		if (sum <= 0) { crb_set_error("tap sum %lld is not positive for phase row %u (the reference would divide by it, H:1025)", (long long)sum, r); goto fail; }
		recip = (int64_t)0x80000000ll / sum;
		if (recip > 0x7FFFFFFF) { crb_set_error("normaliser does not fit 32 bits for phase row %u", r); goto fail; }
		/* each chain (signed columns ride the positive one): |acc| <= 32768 * sum|k| / 65536 must stay below 2^31 */
		if (sum_pos + sum_neg >= ((int64_t)1 << 32)) { crb_set_error("accumulator could overflow 32 bits for phase row %u", r); goto fail; }
		/* |acc_pos - acc_neg| * recip fits int64 trivially; the final sample must fit int32 */
		if (((sum_pos + sum_neg) / 2 + 1) * recip / 32768 >= ((int64_t)1 << 31)) { crb_set_error("output could overflow 32 bits for phase row %u", r); goto fail; }
		/* one-instruction normalisers (crb_device.cu normalise()): mode 2 needs |recip - 32768| < 16384,
		   mode 1 needs recip < 65536 and acc << 1 to fit */
		if (g->norm_mode == 3 && (sum_pos + sum_neg) / 2 >= ((int64_t)1 << 17)) g->norm_mode = 2;   /* acc no longer its own bias */
		/* the unstretched kernel's three-instruction normaliser (crb_kernels.cuh normalise_t16): acc * 2 * (recip - 32768) + 65535 in 32 bits */
		if (((sum_pos + sum_neg) / 2 + 1) * 2 * (recip > 32768 ? recip - 32768 : 32768 - recip) + 65535 >= ((int64_t)1 << 31)) norm_t16_ok = 0;
		if (g->norm_mode >= 2 && !(recip > 16384 && recip < 49152)) g->norm_mode = 1;
		if (g->norm_mode == 1 && !(recip < 65536 && sum_pos + sum_neg < ((int64_t)1 << 30))) g->norm_mode = 0;
		row[n_cols] = (int32_t)recip;
	}
	if (g->norm_mode)
		for (r = 0; r < n_rows; ++r) {
			int32_t *word = &plan->host_rows[(size_t)r * g->row_words + n_cols];
			*word = (int32_t)(((int64_t)*word - 32768) * (g->norm_mode >= 2 ? 131072 : 65536));
		}

	/* 6. a breakpoint whose two rows came out identical (the extra tap was zero-weight and got
	      dropped) is not a breakpoint: merge, so that e.g. the unstretched kernel is row = e >> 6 */
	for (i = 0; i < g->n_breaks;) {
		uint32_t rb = 0;
		while (rb < n_rows && row_of_e_start[rb] != g->breaks[i]) ++rb;
		if (rb > 0 && rb < n_rows
		    && memcmp(plan->host_rows + (size_t)rb * g->row_words, plan->host_rows + (size_t)(rb - 1) * g->row_words, g->row_words * sizeof(int32_t)) == 0) {
			memmove(plan->host_rows + (size_t)rb * g->row_words, plan->host_rows + (size_t)(rb + 1) * g->row_words, (size_t)(n_rows - rb - 1) * g->row_words * sizeof(int32_t));
			memmove(row_of_e_start + rb, row_of_e_start + rb + 1, (n_rows - rb - 1) * sizeof *row_of_e_start);
			--n_rows;
			memmove(g->breaks + i, g->breaks + i + 1, (g->n_breaks - i - 1) * sizeof g->breaks[0]);
			--g->n_breaks;
		} else {
			++i;
		}
	}
	g->n_rows = n_rows;
	for (i = g->n_breaks; i < CRB_MAX_BREAKS; ++i) g->breaks[i] = 0xFFFFFFFFu;   /* never reached: the kernel compares all four */

	/* 7. prove the device's row formula for every phase */
	{
		uint32_t row = 0;
		for (e = 0; e < 65536; ++e) {
			uint32_t dev_row = (uint32_t)((step * (e + delta)) >> 16) - g->ks0;
			for (i = 0; i < g->n_breaks; ++i) dev_row += (e >= g->breaks[i]);
			while (row + 1 < n_rows && row_of_e_start[row + 1] <= e) ++row;
			if (dev_row != row) { crb_set_error("internal: device row formula mismatch at phase %u (%u vs %u)", e, dev_row, row); goto fail; }
		}
	}

	g->channels = channels;
	g->increment = (uint32_t)increment;
	g->step = (uint32_t)step;
	g->delta = (uint32_t)delta;
	g->radius_int = (uint32_t)radius_int;
	g->radius_fx = (uint32_t)radius_fx;
	/* (the mono kernel computes adjacent frame pairs from one window: it needs increment <= 1, which ClownResampler_LowestLevel_Configure
	   guarantees for an unstretched kernel -- H:968 -- but a hand-built state need not) */
	g->unstretched5 = (norm_t16_ok && increment <= CRB_FX_ONE && step == 1024 && delta == 0 && g->n_breaks == 0 && n_rows == 1024 && g->ks0 == 0 && n_cols == 5 && n_runs == 4 && g->norm_mode == 3
		&& g->runs[0].len == 1 && !g->runs[0].negative && !g->runs[0].big && g->runs[1].len == 1 && g->runs[1].negative && !g->runs[1].big
		&& g->runs[2].len == 2 && !g->runs[2].negative && g->runs[2].big && g->runs[3].len == 1 && g->runs[3].negative && !g->runs[3].big
		&& g->runs[0].off == 0 && g->runs[1].off == 1 && g->runs[2].off == 2 && g->runs[3].off == 4);
	g->lane_stride = 1;
	if (g->unstretched5) {
		/* the kernel keeps the chains {tap 3, tap 0}, {tap 2} and {taps 1, 4} in 16-bit fields (crb_kernels.cuh mac_hi16): the
		   weights of a chain must sum to at most 65536 in every phase row, or the plan goes to the general kernel */
		for (r = 0; r < n_rows && g->unstretched5; ++r) {
			const int32_t *row = plan->host_rows + (size_t)r * g->row_words;   /* small columns hold |k| << 16, big ones |k| */
			const int64_t k0 = (uint32_t)row[0] >> 16, k1 = (uint32_t)row[1] >> 16, k2 = row[2], k3 = row[3], k4 = (uint32_t)row[4] >> 16;
			if (k0 + k3 > 65536 || k2 > 65536 || k1 + k4 > 65536) g->unstretched5 = 0;
		}
	}
	if (g->unstretched5) {
		/* the unstretched kernel's five weights and reciprocal pack into 16 bytes (one LDS.128):
		   { k2, k3, (k1 << 16) | k0, (k4 << 16) | (2 * (recip - 32768) & 0xFFFF) }, k0 k1 k4 < 32768 */
		int32_t *packed = (int32_t *)calloc((size_t)n_rows * 4, sizeof(int32_t));
		if (!packed) { crb_set_error("out of host memory"); rc = -5; goto fail; }
		for (r = 0; r < n_rows; ++r) {
			const int32_t *row = plan->host_rows + (size_t)r * g->row_words;
			packed[4 * r + 0] = row[2];
			packed[4 * r + 1] = row[3];
			packed[4 * r + 2] = (int32_t)(((uint32_t)row[1] & 0xFFFF0000u) | ((uint32_t)row[0] >> 16));
			packed[4 * r + 3] = (int32_t)(((uint32_t)row[4] & 0xFFFF0000u) | ((uint32_t)row[5] >> 16));
		}
		free(plan->host_rows);
		plan->host_rows = packed;
		g->row_words = 4;
	}

	/* the slightly stretched kernel has no column rotation: frame strides that pile its sample loads onto a few banks
	   (4 or 8 channels at 2:1, 8 channels at 3:2) stay on the general kernel.  Thresholds from same-box timings of
	   both kernels over 3..8 channels x {48->44.1, 48->32, 96->48} */
	if (!g->unstretched5 && taps_max <= 12 && channels <= 8 && !no_small()
	    && (channels <= 2 || small_kernel_load_wavefronts(increment, channels, taps_max <= 6 ? 6 : (taps_max + 1) & ~1u)
	                         <= ((channels % 2 == 0 && taps_max <= 8) ? 5.0 : 3.0))) {
		/* 7a. slightly stretched kernels (down-sampling by less than about 2): too few taps for the general kernel's
		       column groups to pay off.  Rows hold the signed weights in tap order (the "signed big" form of
		       crb_device.cu: multiplicand sample << 16, bias sample ^ (k >> 31)), then the reciprocal word. */
		const uint32_t taps = taps_max <= 6 ? 6 : (taps_max + 1) & ~1u;
		const uint32_t rw = (taps + 1 + 3) & ~3u;
		int32_t *sk = (int32_t *)calloc((size_t)n_rows * rw, sizeof(int32_t));
		if (!sk) { crb_set_error("out of host memory"); rc = -5; goto fail; }
		for (r = 0; r < n_rows; ++r) {
			const phase_key key = by_e[row_of_e_start[r]];
			for (i = 0; i < key.ntaps; ++i) sk[(size_t)r * rw + i] = plan->host_table[key.ks + i * step];
			sk[(size_t)r * rw + taps] = plan->host_rows[(size_t)r * g->row_words + n_cols];
		}
		free(plan->host_rows);
		plan->host_rows = sk;
		g->small_taps = taps;
		g->row_words = rw;
		g->n_cols = n_cols = taps;
		g->n_runs = n_runs = 1;
		g->runs[0].col = 0; g->runs[0].len = (int32_t)taps; g->runs[0].off = 0; g->runs[0].negative = 2; g->runs[0].big = 1;
	}

	if (!g->unstretched5 && !g->small_taps) {
		/* 7b. regroup the columns for the general kernel: runs of the same (sign, form) become adjacent, every
		       group is padded to an even column count so that the kernel fetches two weights and two frame
		       offsets per 64-bit load, and the rows get a stride of 2 mod 4 words (conflict-light 64-bit loads
		       from different rows).  The per-column frame offsets (bytes) follow the rows.
		       The thread -> frame stride and the column rotation come out of the bank model above. */
		uint32_t *old_col[CRB_GROUPS] = { NULL }, *off[CRB_GROUPS] = { NULL };
		uint32_t count[CRB_GROUPS] = { 0 }, first[CRB_GROUPS], kind[CRB_GROUPS] = { 0 }, order[CRB_MAX_RUNS], new_col_of_old[1024];
		uint32_t n_total = 0, key, q, new_words, n_order = 0, widest = 0, best_mask = 0, best_rot = 0, best_shift = 0, best_stride = 1;
		uint32_t ng = 6;                          /* groups in use */
		uint8_t *col_cls = NULL, *col_big = NULL;
		uint32_t *col_tap = NULL, *singles = NULL;
		int64_t *chain_sum[CRB_GROUPS] = { NULL };
		chain_col *sorted = NULL;
		int32_t *regrouped = NULL;
		int ok = n_cols <= 1000, chains = 0, attempt;
		const uint32_t old_words = g->row_words;
#define CRB_PLAIN(r_, oc_) (col_big[oc_] ? (int64_t)plan->host_rows[(size_t)(r_) * old_words + (oc_)] : (int64_t)(plan->host_rows[(size_t)(r_) * old_words + (oc_)] >> 16))
		if (!ok) crb_set_error("kernel too wide for the tiled kernel");
		for (key = 0; key < CRB_GROUPS && ok; ++key) {
			old_col[key] = (uint32_t *)calloc(n_cols + 2, sizeof(uint32_t));
			off[key] = (uint32_t *)calloc(n_cols + 2, sizeof(uint32_t));
			chain_sum[key] = (int64_t *)calloc(n_rows, sizeof(int64_t));
			if (!old_col[key] || !off[key] || !chain_sum[key]) { crb_set_error("out of host memory"); rc = -5; ok = 0; }
		}
		if (ok) {
			col_cls = (uint8_t *)calloc(n_cols, 1); col_big = (uint8_t *)calloc(n_cols, 1);
			col_tap = (uint32_t *)calloc(n_cols, sizeof(uint32_t));
			sorted = (chain_col *)calloc(n_cols, sizeof *sorted);
			singles = (uint32_t *)calloc(n_cols + 2, sizeof(uint32_t));
			if (!col_cls || !col_big || !col_tap || !sorted || !singles) { crb_set_error("out of host memory"); rc = -5; ok = 0; }
		}
		if (ok) {
			/* the chain form multiplies sample and weight in 32 bits: every |k| must be at most 65536 (65535 where the sign of the
			   weight can make the product positive: -32768 * -65536 does not fit) */
			/* measured: the chain form wins 3-5 % with packed sample loads (4, 6, 8 channels) and loses where the
			   shared-memory pipe is the bound (1, 2 channels) or the columns rotate (more, smaller groups) */
			chains = !no_chains() && (channels == 4 || channels == 6 || channels == 8);
			for (q = 0; q < n_runs; ++q)
				for (i = 0; i < (uint32_t)g->runs[q].len; ++i) {
					const uint32_t oc = (uint32_t)g->runs[q].col + i;
					int64_t maxabs = 0;
					col_cls[oc] = (uint8_t)g->runs[q].negative; col_big[oc] = (uint8_t)g->runs[q].big; col_tap[oc] = (uint32_t)g->runs[q].off + i;
					for (r = 0; r < n_rows; ++r) {
						const int64_t v = CRB_PLAIN(r, oc), a = v < 0 ? -v : v;
						if (a > maxabs) maxabs = a;
					}
					sorted[oc].col = oc; sorted[oc].maxabs = maxabs;
					if (maxabs > (col_cls[oc] == 2 ? 65535 : 65536)) chains = 0;
				}
		}
		for (attempt = 0; attempt < 2 && ok; ++attempt) {
		/* first the IMAD.HI form's six groups; if the plan wants the chain form and the bank model asked for no column rotation,
		   the same again with the chain form's groups */
		memset(count, 0, sizeof count);
		widest = 0; n_order = 0; best_mask = 0; best_rot = 0; best_shift = 0; best_stride = 1; ng = 6;
		g->chain_mode = 0;
		if (ok && attempt == 1) {
			/* chain form: per sign class, first-fit-decreasing of the columns into chains whose |k| sum to at most 65535 in EVERY phase
			   row (exact sums, not the sum of the column maxima); what fits nowhere, or alone, is folded column by column */
			/* two packing orders -- largest first (few, full chains) and smallest first (the most columns inside the chains the group
			   limit allows: wide kernels) -- costed in instructions per channel: 3 per chain column, 4 per single column, the fold
			   and the loop set-up of every group */
			uint32_t cls, pass, ascending = 0;
			double pass_cost[2] = { 0, 0 };
			qsort(sorted, n_cols, sizeof *sorted, chain_col_cmp);
			for (pass = 0; pass < 3; ++pass) {
			if (pass == 2 && ascending == 1) break;      /* the second pass already left the better packing in place */
			if (pass == 2) ascending = 0;
			else ascending = pass;
			ng = 0;
			memset(count, 0, sizeof count);
			for (cls = 0; cls < 3; ++cls) {
				const uint32_t g0 = ng;            /* this class's chains are groups [g0, ng) */
				uint32_t n_singles = 0, a, b;
				for (q = 0; q < n_cols; ++q) {
					const uint32_t sq = ascending ? n_cols - 1 - q : q;
					const uint32_t oc = sorted[sq].col;
					uint32_t placed = 0;
					if (col_cls[oc] != cls) continue;
					if (sorted[sq].maxabs <= 65535) {
						for (key = g0; key <= ng && !placed; ++key) {
							if (key == ng) {                 /* open a new chain: at most three per class, which leaves a group for each class's single columns */
								if (ng - g0 >= 3 || ng >= CRB_GROUPS - (3 - cls)) break;
								memset(chain_sum[key], 0, n_rows * sizeof(int64_t));
								count[key] = 0; kind[key] = cls * 2; ++ng;
							}
							placed = 1;
							for (r = 0; r < n_rows && placed; ++r) {
								const int64_t v = CRB_PLAIN(r, oc);
								if (chain_sum[key][r] + (v < 0 ? -v : v) > 65535) placed = 0;
							}
							if (placed) {
								for (r = 0; r < n_rows; ++r) { const int64_t v = CRB_PLAIN(r, oc); chain_sum[key][r] += v < 0 ? -v : v; }
								old_col[key][count[key]++] = oc;
							}
						}
					}
					if (!placed) singles[n_singles++] = oc;
				}
				/* a chain of one column is a single column */
				for (key = g0; key < ng;)
					if (count[key] == 1) {
						uint32_t *t_col = old_col[key]; int64_t *t_sum = chain_sum[key];
						singles[n_singles++] = old_col[key][0];
						for (a = key; a + 1 < ng; ++a) { old_col[a] = old_col[a + 1]; chain_sum[a] = chain_sum[a + 1]; count[a] = count[a + 1]; kind[a] = kind[a + 1]; }
						old_col[ng - 1] = t_col; chain_sum[ng - 1] = t_sum; count[ng - 1] = 0;
						--ng;
					} else {
						++key;
					}
				/* every group is padded to an even column count: where two groups are odd, move a column instead (the smallest of
				   a chain is its last) */
				for (a = g0; a < ng; ++a) {
					if (!(count[a] & 1u)) continue;
					if (n_singles & 1u) {
						singles[n_singles++] = old_col[a][--count[a]];
						continue;
					}
					for (b = a + 1; b < ng && !(count[b] & 1u); ++b) {}
					if (b < ng) {
						const uint32_t oc = old_col[a][count[a] - 1];
						int fits = 1;
						for (r = 0; r < n_rows && fits; ++r) { const int64_t v = CRB_PLAIN(r, oc); if (chain_sum[b][r] + (v < 0 ? -v : v) > 65535) fits = 0; }
						--count[a];
						if (fits) {
							for (r = 0; r < n_rows; ++r) { const int64_t v = CRB_PLAIN(r, oc); chain_sum[b][r] += v < 0 ? -v : v; }
							old_col[b][count[b]++] = oc;
						} else {
							singles[n_singles++] = oc;
							singles[n_singles++] = old_col[b][--count[b]];
						}
					}
				}
				if (n_singles) {
					memcpy(old_col[ng], singles, n_singles * sizeof(uint32_t));
					count[ng] = n_singles; kind[ng] = cls * 2 + 1;
					++ng;
				}
			}
			if (pass < 2) {
				for (key = 0; key < ng; ++key)
					pass_cost[pass] += ((kind[key] & 1u) ? 4.0 : 3.0) * (double)((count[key] + 1u) & ~1u) + 2.5;
				if (pass == 1) ascending = pass_cost[1] < pass_cost[0];
			}
			}
			for (key = 0; key < ng; ++key) {
				for (i = 0; i < count[key]; ++i) off[key][i] = col_tap[old_col[key][i]] * 2u * channels;
				if (count[key] & 1u) {
					old_col[key][count[key]] = 0xFFFFFFFFu;
					off[key][count[key]] = off[key][count[key] - 1];
					++count[key];
				}
				if (count[key] > count[widest]) widest = key;
			}
			g->chain_mode = 1;
		}
		for (key = 0; key < CRB_GROUPS && ok && !g->chain_mode; ++key) {
			kind[key] = key;
			if (key >= 6) continue;
			for (q = 0; q < n_runs; ++q)
				if ((uint32_t)(g->runs[q].negative * 2 + g->runs[q].big) == key) {
					for (i = 0; i < (uint32_t)g->runs[q].len; ++i) {
						old_col[key][count[key]] = (uint32_t)g->runs[q].col + i;
						off[key][count[key]] = (uint32_t)(g->runs[q].off + i) * 2u * channels;
						++count[key];
					}
					order[n_order++] = q;
				}
			if (count[key] & 1u) {                /* zero-weight pad column reading a frame that is read anyway */
				old_col[key][count[key]] = 0xFFFFFFFFu;
				off[key][count[key]] = off[key][count[key] - 1];
				++count[key];
			}
			if (count[key] > count[widest]) widest = key;
		}
		if (ok) {
			/* candidates: (stride, no rotation) as before, then rotations for stride 1 and for the best stride */
			bank_model m;
			double best_cost, base_cost;
			uint32_t stride, mask, rot, shift, pass;
			m.increment = increment; m.channels = channels; m.threads = CRB_NT(channels);
			m.step = (uint32_t)step; m.delta = (uint32_t)delta; m.ks0 = g->ks0; m.n_breaks = g->n_breaks; m.breaks = g->breaks;
			m.col_off = off[widest]; m.count = count[widest];
			m.first = 0;
			for (key = 0; key < widest; ++key) m.first += count[key];
			new_words = 1;
			for (key = 0; key < CRB_GROUPS; ++key) new_words += count[key];
			while ((new_words & 3u) != 2u) ++new_words;
			m.row_words = new_words;
			best_cost = iteration_wavefronts(&m, 1, 0, 0, 0);
			for (stride = 3; stride < 64; stride += 2) {
				const double cost = iteration_wavefronts(&m, stride, 0, 0, 0);
				if (cost < best_cost * 0.97) { best_cost = cost; best_stride = stride; }
			}
			base_cost = best_cost;
			{
				const uint32_t plain_stride = best_stride;
				for (mask = 1; mask <= 31 && mask < count[widest] / 2 && !g->chain_mode; mask = mask * 2 + 1) {
					/* the copies behind the rotating groups widen the rows */
					uint32_t words = 1, first_w = 0;
					for (key = 0; key < CRB_GROUPS; ++key) {
						if (key == widest) first_w = words - 1;
						words += count[key] + (count[key] / 2 > mask ? 2 * mask : 0);
					}
					while ((words & 3u) != 2u) ++words;
					m.row_words = words;
					m.first = first_w;
					for (pass = 0; pass < 2; ++pass) {
						stride = pass ? plain_stride : 1;
						if (pass && plain_stride == 1) break;
						for (shift = 0; shift < 5; ++shift)
							for (rot = 1; rot <= 7 && (rot == 1 || rot <= mask); rot += 2) {
								const double cost = iteration_wavefronts(&m, stride, rot, shift, mask);
								/* rotation costs table space (smaller tiles): take it only for a clear gain over the unrotated
								   layout, and a wider mask only for a clear gain over a narrower one */
								if (cost < best_cost * (mask > best_mask && best_mask ? 0.85 : 0.97) && cost < base_cost * 0.75) {
									best_cost = cost; best_mask = mask; best_rot = rot; best_shift = shift; best_stride = stride;
								}
							}
					}
				}
			}
		}
		if (attempt == 0 && chains && best_rot == 0) continue;
		break;
		}
		if (ok) {
			g->lane_stride = best_stride;
			g->rot = best_rot;
			g->rot_shift = best_shift;
			g->rot_mask = best_mask;
			for (key = 0; key < CRB_GROUPS; ++key) {
				const uint32_t rotates = best_rot && count[key] / 2 > best_mask;
				first[key] = n_total;
				g->group_rot[key] = rotates ? 0xFFFFFFFFu : 0u;
				n_total += count[key] + (rotates ? 2 * best_mask : 0);
			}
			new_words = n_total + 1;
			while ((new_words & 3u) != 2u) ++new_words;
			regrouped = (int32_t *)calloc((size_t)n_rows * new_words + n_total + 4, sizeof(int32_t));
			if (!regrouped) { crb_set_error("out of host memory"); rc = -5; ok = 0; }
		}
		if (ok) {
			int32_t *col_off = regrouped + (size_t)n_rows * new_words;
			for (key = 0; key < CRB_GROUPS; ++key) {
				const uint32_t copies = g->group_rot[key] ? 2 * best_mask : 0;
				for (i = 0; i < count[key] + copies; ++i) {
					const uint32_t src = i < count[key] ? i : i - count[key];
					const uint32_t oc = old_col[key][src];
					col_off[first[key] + i] = (int32_t)off[key][src];
					if (i < count[key] && oc != 0xFFFFFFFFu) new_col_of_old[oc] = first[key] + i;
					for (r = 0; r < n_rows; ++r)
						regrouped[(size_t)r * new_words + first[key] + i] = oc == 0xFFFFFFFFu ? 0
							: g->chain_mode ? (int32_t)CRB_PLAIN(r, oc) : plan->host_rows[(size_t)r * g->row_words + oc];
				}
				g->groups[key][0] = first[key];
				g->groups[key][1] = count[key];
				g->group_kind[key] = (uint8_t)kind[key];
			}
			g->n_groups = ng;
			for (r = 0; r < n_rows; ++r)
				regrouped[(size_t)r * new_words + n_total] = plan->host_rows[(size_t)r * g->row_words + n_cols];
			if (g->chain_mode) {
				g->n_runs = 0;     /* the groups and the per-column offsets describe a chain-form table; runs of consecutive taps do not */
			} else {   /* the runs keep describing the (moved) columns for the tests' arithmetic model */
				crb_run moved[CRB_MAX_RUNS];
				for (q = 0; q < n_order; ++q) { moved[q] = g->runs[order[q]]; moved[q].col = (int32_t)new_col_of_old[g->runs[order[q]].col]; }
				memcpy(g->runs, moved, n_order * sizeof moved[0]);
			}
			g->const_offsets = 0;
			/* (measured again with the chain form on 8 channels: 5 % slower through the constant cache) */
			if (!best_rot && (channels & 1u) && channels < 8 && n_total <= CRB_CONST_COLS && taps_max * 2u * channels < 65536u) {
				for (i = 0; i < n_total; ++i) g->col_off16[i] = (uint16_t)col_off[i];
				g->const_offsets = 1;
			}
			free(plan->host_rows);
			plan->host_rows = regrouped;
			g->row_words = new_words;
			g->n_cols = n_total;
			g->colinfo_words = n_total;
			n_cols = n_total;
		}
		for (key = 0; key < CRB_GROUPS; ++key) { free(old_col[key]); free(off[key]); free(chain_sum[key]); }
		free(col_cls); free(col_big); free(col_tap); free(sorted); free(singles);
#undef CRB_PLAIN
		if (!ok) goto fail;
	}

	/* 8. tile geometry: a ring of CRB_RING_STAGES input windows next to the table.  Prefer four CTAs
	      per SM with big tiles, then two, then one; the direct kernel is the last resort. */
	{
		/* resident CTAs per SM each kernel instantiation is compiled for (crb_device.cu launch bounds):
		   see CRB_NT / CRB_CTAS in crb_internal.h */
		const uint32_t want = CRB_CTAS_K(channels, g->unstretched5 ? 1u : g->small_taps);
		const uint32_t budgets[3] = { 227 * 1024 / want - 1024, 112 * 1024, 0 };
		const uint32_t nt = CRB_NT_K(channels, g->unstretched5);
		const uint32_t min_tile[3] = { channels == 8 ? nt : 2u * nt, nt, 32 };
		const uint32_t frame_bytes = 2 * channels;
		const uint32_t rows_bytes = ((n_rows * g->row_words + g->colinfo_words) * 4 + 15u) & ~15u;
		uint32_t tile_out, b;
		plan->kernel_kind = 1;
		g->n_stages = CRB_RING_STAGES;
		for (b = 0; b < 3 && plan->kernel_kind == 1; ++b) {
			uint32_t budget = budgets[b] ? budgets[b] : smem_budget_bytes;
			if (budget > smem_budget_bytes) budget = smem_budget_bytes;
			for (tile_out = CRB_FULL_TILE_K(channels, g->unstretched5); tile_out >= min_tile[b]; tile_out >>= 1) {
				const uint64_t span = ((uint64_t)tile_out * increment + 65535) / 65536; /* frames between first and last window start, rounded up */
				const uint64_t in_frames = span + taps_max + 2 + 16;                    /* + widest window + start rounding + alignment slack */
				uint64_t stage = ((in_frames * frame_bytes + 15) & ~(uint64_t)15) + 16;
				uint64_t slot[3] = { 0, 0, 0 };
				uint32_t l;
				if ((uint64_t)tile_out * increment + ((uint64_t)20 << 16) >= ((uint64_t)1 << 31)) continue; /* 32-bit tile-relative positions */
				if (rows_bytes + CRB_RING_STAGES * stage + CRB_CTRL_BYTES > budget) continue;
				/* unstretched kernel: a tile may instead hold tile_out / 2 (/ 4) frames of each of two (four) lockstep streams;
				   every stream's window carries its own halo and alignment slack, so the stage grows a little -- if the budget allows */
				slot[0] = stage;
				for (l = 1; l < 3 && g->unstretched5 && (tile_out >> l) >= 32; ++l) {
					const uint64_t span_l = ((uint64_t)(tile_out >> l) * increment + 65535) / 65536;
					const uint64_t bytes_l = (((span_l + taps_max + 2 + 16) * frame_bytes + 15) & ~(uint64_t)15) + 16;
					const uint64_t grown = (bytes_l << l) > stage ? (bytes_l << l) : stage;
					if (rows_bytes + CRB_RING_STAGES * grown + CRB_CTRL_BYTES > budget) break;
					slot[l] = bytes_l;
					stage = grown;
				}
				g->tile_out = tile_out;
				g->tile_in_frames = (uint32_t)in_frames;
				g->stage_bytes = (uint32_t)stage;
				g->lock_slot_bytes[0] = (uint32_t)slot[0]; g->lock_slot_bytes[1] = (uint32_t)slot[1]; g->lock_slot_bytes[2] = (uint32_t)slot[2];
				plan->smem_bytes = (uint32_t)(rows_bytes + CRB_RING_STAGES * stage + CRB_CTRL_BYTES);
				plan->kernel_kind = 0;
				break;
			}
		}
		if (force_direct()) plan->kernel_kind = 1;
		if (plan->kernel_kind == 1) {
			g->tile_out = CRB_DIRECT_THREADS;
			plan->smem_bytes = 0;
		}
	}

	plan->cfg_radius_fx = (uint32_t)radius_fx;
	plan->cfg_radius_int = (uint32_t)radius_int;
	plan->cfg_delta = (uint32_t)delta;
	plan->cfg_step = (uint32_t)step;
	plan->table_hash = crb_hash_table(table);
	rc = 0;
fail:
	free(by_e); free(row_key); free(row_of_e_start); free(tap_pos); free(tap_neg); free(col_of_tap_pos); free(col_of_tap_neg);
	if (rc != 0) {
		free(plan->host_rows); plan->host_rows = NULL;
		free(plan->host_table); plan->host_table = NULL;
	}
	return rc;
}
