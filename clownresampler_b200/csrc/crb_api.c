/*
 * crb_api.c -- plain-C host layer of libclownresampler_b200.so: the reference's C89 API
 * (include/clownresampler.h) and the bulk extensions (include/clownresampler_b200.h) on top of
 * the CUDA layer in crb_device.cu.  No arithmetic of the hot path happens in this file: it
 * computes configurations and closed-form positions (cheap integer code that defines the
 * geometry), moves buffers, launches kernels and delivers frames to callbacks.
 *
 * H = /root/reference/clownresampler.h.
 */
#include "../../include/clownresampler_b200.h"
#include "crb_internal.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef char crb_assert_precomputed[sizeof(ClownResampler_Precomputed) == 49152 ? 1 : -1];
typedef char crb_assert_config[sizeof(ClownResampler_LowestLevel_Configuration) == 32 ? 1 : -1];
typedef char crb_assert_lowlevel[sizeof(ClownResampler_LowLevel_State) == 64 ? 1 : -1];
typedef char crb_assert_highlevel[sizeof(ClownResampler_HighLevel_State) == 8296 ? 1 : -1];

#define FX 65536ul

/* =========================================================================================
 * process-wide context: plan cache and staging lanes of the callback and host-bulk paths.
 *
 * G.lock guards only the tables below (cache lookup, lane hand-out, kept-frame table); it is never held while a
 * kernel runs, while the calling thread waits for the GPU, or while a user callback runs.  Each call that stages
 * host memory borrows a LANE (a stream and SLOTS pinned + device buffer pairs) for its duration, so calls on
 * different states from different threads run concurrently, and an output callback may itself drive another
 * resampler (it borrows another lane).  Devices: see crb_device.cu -- every object remembers its device.
 * ========================================================================================= */
#define PLAN_CACHE 32
#define LANES 8
#ifndef CRB_SLOTS
#define CRB_SLOTS 3
#endif
#define SLOTS CRB_SLOTS

typedef struct crb_slot {
	void *stream, *done;
	void *pin_in, *dev_in, *pin_out, *dev_out;
	size_t in_cap, out_cap;
} crb_slot;

typedef struct crb_lane {
	int busy, device;            /* device: -1 = holds no resources yet */
	crb_slot slots[SLOTS];
} crb_lane;

static struct {
	pthread_mutex_t lock;
	pthread_cond_t lane_free;
	struct ClownResamplerB200_Plan *plans[PLAN_CACHE];
	unsigned long plan_age[PLAN_CACHE], clock;
	crb_lane lanes[LANES];
	int lanes_ready;
} G = { PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, {0}, {0}, 0, {{0}}, 0 };

static void report(const char *where)
{
	/* The reference's API has no error channel; never continue silently. */
	fprintf(stderr, "clownresampler_b200: %s: %s\n", where, ClownResamplerB200_GetLastError());
}

/* the device calls without a handle work on (the reference's API, the allocation helpers), made usable; < 0: none */
static int default_device(void)
{
	const int d = crb_dev_default();
	return d >= 0 ? d : crb_dev_init(-1, 0);
}

int ClownResamplerB200_Init(int device)
{
	return crb_dev_init(device, 1) >= 0 ? CRB200_OK : CRB200_E_NO_DEVICE;
}

int ClownResamplerB200_DeviceCount(void)
{
	return crb_dev_count();
}

static void slot_release(crb_slot *s)
{
	crb_dev_pinned_free(s->pin_in); crb_dev_pinned_free(s->pin_out);
	crb_dev_free(s->dev_in); crb_dev_free(s->dev_out);
	crb_dev_event_destroy(s->done); crb_dev_stream_destroy(s->stream);
	memset(s, 0, sizeof *s);
}

static void lane_drop_resources(crb_lane *lane)
{
	int i;
	if (lane->device >= 0) {
		const int prev = crb_dev_push(lane->device);
		for (i = 0; i < SLOTS; ++i) slot_release(&lane->slots[i]);
		crb_dev_pop(prev);
	}
	lane->device = -1;
}

/* borrows a staging lane on `device` (waits for one when all are busy) */
static crb_lane *lane_acquire(int device)
{
	crb_lane *lane = NULL;
	int i;
	pthread_mutex_lock(&G.lock);
	if (!G.lanes_ready) { for (i = 0; i < LANES; ++i) G.lanes[i].device = -1; G.lanes_ready = 1; }
	for (;;) {
		for (i = 0; i < LANES && !lane; ++i) if (!G.lanes[i].busy && G.lanes[i].device == device) lane = &G.lanes[i];
		for (i = 0; i < LANES && !lane; ++i) if (!G.lanes[i].busy && G.lanes[i].device < 0) lane = &G.lanes[i];
		for (i = 0; i < LANES && !lane; ++i) if (!G.lanes[i].busy) lane = &G.lanes[i];
		if (lane) break;
		pthread_cond_wait(&G.lane_free, &G.lock);
	}
	lane->busy = 1;
	pthread_mutex_unlock(&G.lock);
	if (lane->device != device) {          /* a lane that served another device: its buffers live there */
		lane_drop_resources(lane);
		lane->device = device;
	}
	return lane;
}

static void lane_release(crb_lane *lane)
{
	pthread_mutex_lock(&G.lock);
	lane->busy = 0;
	pthread_cond_signal(&G.lane_free);
	pthread_mutex_unlock(&G.lock);
}

static void plan_free(struct ClownResamplerB200_Plan *plan)
{
	if (!plan) return;
	crb_dev_plan_release(plan);
	free(plan->host_rows);
	free(plan->host_table);
	free(plan);
}

/* drops one reference (the cache's, a running call's or the owner's); G.lock held */
static void plan_unref_locked(struct ClownResamplerB200_Plan *plan)
{
	if (plan && --plan->refcount == 0) plan_free(plan);
}

static void memo_release_all(void);

void ClownResamplerB200_Shutdown(void)
{
	int i;
	pthread_mutex_lock(&G.lock);
	for (i = 0; i < PLAN_CACHE; ++i) { plan_unref_locked(G.plans[i]); G.plans[i] = NULL; }
	for (i = 0; i < LANES; ++i) if (!G.lanes[i].busy && G.lanes_ready) lane_drop_resources(&G.lanes[i]);
	memo_release_all();
	pthread_mutex_unlock(&G.lock);
}

/* =========================================================================================
 * table, ratio, configuration: host integer / libm code, bit-identical to the reference
 * ========================================================================================= */

/* Lanczos-3 window (H:892-908). */
static double crb_lanczos(double x)
{
	static const double pi = 3.1415926535897932384626433832795028841971693993751058209749445923078164062862089986280348253421170679;
	const double px = x * pi, pxr = px / (double)CLOWNRESAMPLER_KERNEL_RADIUS;
	return x == 0.0 ? 1.0 : (sin(px) * sin(pxr)) / (px * pxr);
}

void ClownResampler_Precompute(ClownResampler_Precomputed *precomputed)
{
	/* H:955-961: same expression order, same libm, truncating cast */
	const size_t n = sizeof precomputed->lanczos_kernel_table / sizeof precomputed->lanczos_kernel_table[0];
	size_t i;
	for (i = 0; i < n; ++i)
		precomputed->lanczos_kernel_table[i] = (cc_s32l)(crb_lanczos(((double)i / (double)n * 2.0 - 1.0) * (double)CLOWNRESAMPLER_KERNEL_RADIUS) * (double)FX);
}

/* floor(a * 65536 / b) digit by digit in base 65536, with the reference's sentinels (H:913-953). */
static cc_u32f crb_ratio(cc_u32f a, cc_u32f b)
{
	cc_u32f digit[3], quot[3], carry = 0, value;
	int i;
	if (a == 0 || b == 0)
		return 0xFFFFFFFF;
	digit[0] = a / FX; digit[1] = a % FX; digit[2] = 0;
	for (i = 0; i < 3; ++i) {
		const cc_u32f v = digit[i] | carry * FX;
		quot[i] = v / b;
		carry = v % b;
	}
	if (quot[0] != 0 || quot[1] >= FX)
		return 0xFFFFFFFF;
	value = quot[1] * FX + quot[2];
	return value ? value : 1;
}

cc_bool ClownResampler_LowestLevel_Configure(ClownResampler_LowestLevel_Configuration *configuration,
	cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate)
{
	/* H:963-984 */
	cc_u32f low = input_sample_rate, scale, inverse;
	if (output_sample_rate < low) low = output_sample_rate;
	if (low_pass_filter_sample_rate < low) low = low_pass_filter_sample_rate;
	scale = crb_ratio(input_sample_rate, low);
	inverse = crb_ratio(low, input_sample_rate);
	if (scale >= 0x1000 * FX)
		return cc_false;
	configuration->stretched_kernel_radius = CLOWNRESAMPLER_KERNEL_RADIUS * scale;
	configuration->integer_stretched_kernel_radius = (configuration->stretched_kernel_radius + (FX - 1)) / FX;
	configuration->stretched_kernel_radius_delta = configuration->integer_stretched_kernel_radius * FX - configuration->stretched_kernel_radius;
	configuration->kernel_step_size = (size_t)((long)CLOWNRESAMPLER_KERNEL_RESOLUTION * (long)inverse / (long)FX);
	return cc_true;
}

cc_bool ClownResampler_LowLevel_Adjust(ClownResampler_LowLevel_State *resampler,
	cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate)
{
	resampler->increment = crb_ratio(input_sample_rate, output_sample_rate);          /* H:1054 */
	return ClownResampler_LowestLevel_Configure(&resampler->lowest_level, input_sample_rate, output_sample_rate, low_pass_filter_sample_rate);
}

cc_bool ClownResampler_LowLevel_Init(ClownResampler_LowLevel_State *resampler, cc_u8f channels,
	cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate)
{
	resampler->channels = channels;                                                    /* H:1046-1049 */
	resampler->position_integer = 0;
	resampler->position_fractional = 0;
	return ClownResampler_LowLevel_Adjust(resampler, input_sample_rate, output_sample_rate, low_pass_filter_sample_rate);
}

/* =========================================================================================
 * closed forms of the position generator
 * ========================================================================================= */
typedef unsigned __int128 u128;

static u128 position_of(const ClownResampler_LowLevel_State *st, u128 n)
{
	return ((u128)st->position_integer << 16) + st->position_fractional + n * (u128)st->increment;
}

size_t ClownResamplerB200_CountOutputFrames(const ClownResampler_LowLevel_State *state, size_t total_input_frames)
{
	const u128 start = position_of(state, 0), end = (u128)total_input_frames << 16;
	if (start >= end || state->increment == 0)
		return 0;
	return (size_t)((end - start + state->increment - 1) / state->increment);
}

void ClownResamplerB200_AdvanceState(ClownResampler_LowLevel_State *state, size_t *total_input_frames,
	size_t frames_emitted, int stopped)
{
	const u128 p = position_of(state, frames_emitted);
	const size_t whole = (size_t)(p >> 16);
	state->position_fractional = (cc_u32f)(p & 0xFFFF);
	if (stopped) {                       /* H:1084-1088 */
		const size_t used = whole < *total_input_frames ? whole : *total_input_frames;
		*total_input_frames -= used;
		state->position_integer = whole - used;
	} else {                             /* H:1063-1067 */
		state->position_integer = whole - *total_input_frames;
		*total_input_frames = 0;
	}
}

int ClownResamplerB200_SegmentStream(const ClownResampler_LowLevel_State *state, size_t total_input_frames,
	size_t segment_count, size_t index, size_t *first_output_frame, size_t *output_frames,
	size_t *first_padded_input_frame, size_t *padded_input_frames, size_t *position_integer, cc_u32f *position_fractional)
{
	const size_t n_total = ClownResamplerB200_CountOutputFrames(state, total_input_frames);
	const size_t R = state->lowest_level.integer_stretched_kernel_radius;
	size_t n0, n1, first_in, last_in;
	u128 p0, p1;
	if (segment_count == 0 || index >= segment_count) { crb_set_error("segment %zu of %zu", index, segment_count); return CRB200_E_ARGUMENT; }
	n0 = (size_t)((u128)n_total * index / segment_count);
	n1 = (size_t)((u128)n_total * (index + 1) / segment_count);
	*first_output_frame = n0;
	*output_frames = n1 - n0;
	if (n1 == n0) { *first_padded_input_frame = 0; *padded_input_frames = 0; *position_integer = 0; *position_fractional = 0; return CRB200_OK; }
	p0 = position_of(state, n0);
	p1 = position_of(state, n1 - 1);
	first_in = (size_t)(p0 >> 16);                    /* padded-buffer frame of the first window's base (H:995: pos + min_rel >= pos) */
	/* the last window ends at most at pos + 2R (H:996: pos + R + max_rel); one frame more makes the slice a
	   well-formed buffer of (pos - first + 1) frames plus 2R of padding, so that H:1063 does not stop early */
	last_in = (size_t)(p1 >> 16) + 2 * R + 1;
	if (last_in > total_input_frames + 2 * R) last_in = total_input_frames + 2 * R;
	*first_padded_input_frame = first_in;
	*padded_input_frames = last_in - first_in;
	*position_integer = 0;
	*position_fractional = (cc_u32f)(p0 & 0xFFFF);
	return CRB200_OK;
}

/* =========================================================================================
 * plans
 * ========================================================================================= */
static unsigned long g_plans_built;     /* diagnostics: ClownResamplerB200_PlansBuilt */

/* builds and uploads a plan on `device` for the geometry and channel count of `st`, tiles sized for `increment` */
static struct ClownResamplerB200_Plan *plan_create(const ClownResampler_Precomputed *pre, const ClownResampler_LowLevel_State *st, cc_u32f increment, int device)
{
	struct ClownResamplerB200_Plan *plan;
	int prev;
	plan = (struct ClownResamplerB200_Plan *)calloc(1, sizeof *plan);
	if (!plan) { crb_set_error("out of host memory"); return NULL; }
	prev = crb_dev_push(device);
	if (crb_plan_build_host(plan, pre->lanczos_kernel_table, st->lowest_level.stretched_kernel_radius,
	                        st->lowest_level.integer_stretched_kernel_radius, st->lowest_level.stretched_kernel_radius_delta,
	                        st->lowest_level.kernel_step_size, increment, st->channels, crb_dev_smem_optin(device)) != 0
	    || crb_dev_plan_upload(plan) != 0) {
		crb_dev_pop(prev);
		plan_free(plan);
		return NULL;
	}
	crb_dev_pop(prev);
	plan->refcount = 1;
	__atomic_add_fetch(&g_plans_built, 1, __ATOMIC_RELAXED);
	return plan;
}

unsigned long ClownResamplerB200_PlansBuilt(void)
{
	return __atomic_load_n(&g_plans_built, __ATOMIC_RELAXED);
}

ClownResamplerB200_Plan *ClownResamplerB200_PlanCreate(const ClownResampler_Precomputed *precomputed, const ClownResampler_LowLevel_State *state)
{
	int device;
	if (!precomputed || !state) { crb_set_error("null argument"); return NULL; }
	if ((device = default_device()) < 0) return NULL;
	return plan_create(precomputed, state, state->increment, device);
}

ClownResamplerB200_Plan *ClownResamplerB200_PlanCreateOnDevice(const ClownResampler_Precomputed *precomputed, const ClownResampler_LowLevel_State *state, int device)
{
	if (!precomputed || !state) { crb_set_error("null argument"); return NULL; }
	if ((device = crb_dev_init(device, 0)) < 0) return NULL;
	return plan_create(precomputed, state, state->increment, device);
}

void ClownResamplerB200_PlanDestroy(ClownResamplerB200_Plan *plan)
{
	pthread_mutex_lock(&G.lock);
	plan_unref_locked(plan);
	pthread_mutex_unlock(&G.lock);
}

int ClownResamplerB200_PlanGetInfo(const ClownResamplerB200_Plan *plan, ClownResamplerB200_PlanInfo *info)
{
	if (!plan || !info) { crb_set_error("null argument"); return CRB200_E_ARGUMENT; }
	info->channels = plan->geo.channels;
	info->increment = plan->geo.increment;
	info->phases = plan->geo.n_rows;
	info->taps_max = plan->geo.taps_max;
	info->columns = plan->geo.n_cols;
	info->runs = plan->geo.n_runs;
	info->tile_output_frames = plan->geo.tile_out;
	info->tile_input_frames = plan->geo.tile_in_frames;
	info->smem_bytes = plan->smem_bytes;
	info->kernel_kind = (unsigned)plan->kernel_kind;
	info->mean_taps = plan->mean_taps;
	return CRB200_OK;
}

/* Host-only plan construction for the CPU-side tests (no device needed): serialises the geometry
   as 32-bit words {channels, increment, step, delta, radius_int, radius_fx, ks0, n_breaks,
   breaks[4], n_rows, n_cols, row_words, taps_max, n_runs, tile_out, tile_in_frames, stage_bytes,
   unstretched5, norm_mode, kernel_kind, smem_bytes, lane_stride, rot, rot_shift, rot_mask, groups[12] x {first, columns, rotates, kind}, small_taps,
   chain_mode, n_groups,
   runs[n_runs] x {col, len, off, negative (0, 1, 2 = signed) | big << 2}}
   and copies the rows followed by the per-column frame offsets.  Returns the number of geometry words, or a negative error. */
int ClownResamplerB200_DebugBuildPlanHost(const ClownResampler_Precomputed *pre, const ClownResampler_LowLevel_State *st,
	unsigned smem_budget, unsigned *geometry_words, size_t geometry_capacity, int *rows, size_t rows_capacity)
{
	struct ClownResamplerB200_Plan plan;
	const crb_geometry *g = &plan.geo;
	unsigned head[128];
	size_t n = 0, i;
	int rc;
	memset(&plan, 0, sizeof plan);
	rc = crb_plan_build_host(&plan, pre->lanczos_kernel_table, st->lowest_level.stretched_kernel_radius,
		st->lowest_level.integer_stretched_kernel_radius, st->lowest_level.stretched_kernel_radius_delta,
		st->lowest_level.kernel_step_size, st->increment, st->channels, smem_budget);
	if (rc != 0) return rc;
	head[n++] = g->channels; head[n++] = g->increment; head[n++] = g->step; head[n++] = g->delta;
	head[n++] = g->radius_int; head[n++] = g->radius_fx; head[n++] = g->ks0; head[n++] = g->n_breaks;
	for (i = 0; i < CRB_MAX_BREAKS; ++i) head[n++] = g->breaks[i];
	head[n++] = g->n_rows; head[n++] = g->n_cols; head[n++] = g->row_words; head[n++] = g->taps_max; head[n++] = g->n_runs;
	head[n++] = g->tile_out; head[n++] = g->tile_in_frames; head[n++] = g->stage_bytes; head[n++] = g->unstretched5;
	head[n++] = g->norm_mode; head[n++] = (unsigned)plan.kernel_kind; head[n++] = plan.smem_bytes; head[n++] = g->lane_stride;
	head[n++] = g->rot; head[n++] = g->rot_shift; head[n++] = g->rot_mask;
	for (i = 0; i < CRB_GROUPS; ++i) { head[n++] = g->groups[i][0]; head[n++] = g->groups[i][1]; head[n++] = g->group_rot[i] & 1u; head[n++] = g->group_kind[i]; }
	head[n++] = g->small_taps; head[n++] = g->chain_mode; head[n++] = g->n_groups;
	if (n + 4 * g->n_runs > geometry_capacity || (size_t)g->n_rows * g->row_words + g->colinfo_words > rows_capacity) {
		crb_set_error("debug buffers too small");
		rc = CRB200_E_ARGUMENT;
	} else {
		memcpy(geometry_words, head, n * sizeof head[0]);
		for (i = 0; i < g->n_runs; ++i) {
			geometry_words[n++] = (unsigned)g->runs[i].col; geometry_words[n++] = (unsigned)g->runs[i].len;
			geometry_words[n++] = (unsigned)g->runs[i].off; geometry_words[n++] = (unsigned)(g->runs[i].negative | (g->runs[i].big << 2));
		}
		memcpy(rows, plan.host_rows, ((size_t)g->n_rows * g->row_words + g->colinfo_words) * sizeof(int));
		rc = (int)n;
	}
	free(plan.host_rows);
	free(plan.host_table);
	return rc;
}

/* Identity of a caller's table BY CONTENT: a small registry of the distinct tables seen (normally one: ClownResampler_Precompute
   always produces the same, H:679-681).  Every call compares all 49152 bytes of the table it is handed with the registered copy
   (one memcmp, about a microsecond) -- cheaper than hashing it and exact: a table edited in place, anywhere, gets a new id, so
   neither a cached plan nor frames kept from an earlier call can outlive the contents they were computed from.
   Returns 1..TABLES, or 0 when the registry is full (the call then builds its own plan and keeps no frames). */
#define TABLES 16
static struct { ClownResampler_Precomputed *copy; const ClownResampler_Precomputed *last_seen; } g_tables[TABLES];
static int g_n_tables;

static unsigned table_id(const ClownResampler_Precomputed *pre)
{
	int i, n = __atomic_load_n(&g_n_tables, __ATOMIC_ACQUIRE);
	unsigned id = 0;
	for (i = 0; i < n; ++i)         /* the slot that saw this pointer last first */
		if (g_tables[i].last_seen == pre && memcmp(g_tables[i].copy, pre, sizeof *pre) == 0) return (unsigned)i + 1;
	for (i = 0; i < n; ++i)
		if (memcmp(g_tables[i].copy, pre, sizeof *pre) == 0) { g_tables[i].last_seen = pre; return (unsigned)i + 1; }
	pthread_mutex_lock(&G.lock);
	for (i = n; i < g_n_tables && !id; ++i)    /* registered by another thread meanwhile */
		if (memcmp(g_tables[i].copy, pre, sizeof *pre) == 0) id = (unsigned)i + 1;
	if (!id && g_n_tables < TABLES && (g_tables[g_n_tables].copy = (ClownResampler_Precomputed *)malloc(sizeof *pre)) != NULL) {
		memcpy(g_tables[g_n_tables].copy, pre, sizeof *pre);
		g_tables[g_n_tables].last_seen = pre;
		id = (unsigned)g_n_tables + 1;
		__atomic_store_n(&g_n_tables, g_n_tables + 1, __ATOMIC_RELEASE);
	}
	pthread_mutex_unlock(&G.lock);
	return id;
}

/* The tiles of a cached plan are sized for an increment at or above the caller's: every up-sampling ratio shares the plan
   for increment 1.0, other ratios round up to four significant bits.  The phase table depends on the kernel geometry only
   and every job carries its own 16.16 step, so a stream whose ratio is bent continuously (LowLevel_Adjust / HighLevel_Adjust
   per tick) keeps hitting one plan instead of rebuilding one per ratio. */
static cc_u32f plan_increment_for(cc_u32f increment)
{
	cc_u32f unit = 1;
	if (increment <= FX) return FX;
	while ((increment >> 4) >= unit) unit <<= 1;     /* unit = 2^(floor(log2(increment)) - 3) */
	return (increment + unit - 1) / unit * unit;
}

/* cached plan for the calls without a plan handle, keyed by table contents + kernel geometry + channels + device; returns it
   with one more reference (plan_unref when the call is done), or NULL */
static struct ClownResamplerB200_Plan *plan_cached(const ClownResampler_Precomputed *pre, const ClownResampler_LowLevel_State *st, int device, unsigned table)
{
	struct ClownResamplerB200_Plan *fresh = NULL;
	int i, pass, victim;
	if (st->increment == 0 || st->increment > 0xFFFFFFFFul) { crb_set_error("resampler state has increment %lu; was it initialised with ClownResampler_LowLevel_Init?", st->increment); return NULL; }
	for (pass = 0; pass < 2; ++pass) {
		pthread_mutex_lock(&G.lock);
		for (i = 0; i < PLAN_CACHE; ++i) {
			struct ClownResamplerB200_Plan *p = G.plans[i];
			if (p && table && p->table_id == table && p->geo.channels == st->channels && p->geo.increment >= st->increment
			    && p->cfg_radius_fx == st->lowest_level.stretched_kernel_radius && p->cfg_step == st->lowest_level.kernel_step_size
			    && p->cfg_radius_int == st->lowest_level.integer_stretched_kernel_radius && p->device == device) {
				G.plan_age[i] = ++G.clock;
				++p->refcount;
				pthread_mutex_unlock(&G.lock);
				if (fresh) { pthread_mutex_lock(&G.lock); plan_unref_locked(fresh); pthread_mutex_unlock(&G.lock); }   /* another thread was faster */
				return p;
			}
		}
		if (fresh && !table) {      /* unregistered table: the plan serves this call only */
			pthread_mutex_unlock(&G.lock);
			return fresh;
		}
		if (fresh) {
			victim = 0;
			for (i = 0; i < PLAN_CACHE; ++i) {
				if (!G.plans[i]) { victim = i; break; }
				if (G.plan_age[i] < G.plan_age[victim]) victim = i;
			}
			plan_unref_locked(G.plans[victim]);       /* a call still running on it keeps it alive until it is done */
			G.plans[victim] = fresh;
			G.plan_age[victim] = ++G.clock;
			++fresh->refcount;
			pthread_mutex_unlock(&G.lock);
			return fresh;
		}
		pthread_mutex_unlock(&G.lock);
		/* not cached: build it outside the lock (milliseconds of host work and two device allocations) */
		if (!(fresh = plan_create(pre, st, plan_increment_for(st->increment), device))) return NULL;
		fresh->table_id = table;
	}
	return NULL;
}

static void plan_unref(struct ClownResamplerB200_Plan *plan)
{
	pthread_mutex_lock(&G.lock);
	plan_unref_locked(plan);
	pthread_mutex_unlock(&G.lock);
}

/* =========================================================================================
 * bulk device path
 * ========================================================================================= */
static size_t out_frame_bytes(const struct ClownResamplerB200_Plan *plan, int fmt)
{
	return fmt == CRB200_OUT_S16_CLAMPED ? 2u * plan->geo.channels : fmt == 2 ? 4u * (plan->geo.channels + 1) : 4u * plan->geo.channels;
}

/* CRB200_NO_LOCKSTEP=1 keeps every job on its own (test / A-B hook) */
static int no_lockstep(void)
{
	const char *v = getenv("CRB200_NO_LOCKSTEP");
	return v && v[0] == '1';
}

/* Converts public jobs to device jobs (prefix of tiles included); returns total tiles or -1, and the number of device
   jobs in *n_device.  With an unstretched plan, runs of consecutive jobs that walk through the same phases (same position
   fraction, first frame and frame count -- e.g. a batch of equally long streams) are merged four or two at a time into
   lockstep jobs: the kernel then fetches a frame's phase row once for all of them. */
static int64_t convert_jobs(const struct ClownResamplerB200_Plan *plan, const ClownResamplerB200_Job *jobs, size_t n, crb_device_job *out, size_t *n_device)
{
	const size_t R = plan->geo.radius_int;
	uint64_t tiles = 0;
	size_t i, n_dev = 0;
	for (i = 0; i < n; ++i) {
		const ClownResamplerB200_Job *j = &jobs[i];
		ClownResampler_LowLevel_State st;
		size_t available;
		memset(&st, 0, sizeof st);
		st.position_integer = j->position_integer;
		st.position_fractional = j->position_fractional;
		st.increment = plan->geo.increment;
		available = ClownResamplerB200_CountOutputFrames(&st, j->total_input_frames);
		if (j->position_fractional >= FX) { crb_set_error("job %zu: position_fractional %lu is not a 16-bit fraction", i, j->position_fractional); return -1; }
		if (j->first_output_frame > available || j->output_frames > available - j->first_output_frame) {
			crb_set_error("job %zu asks for output frames [%zu, %zu) but only %zu exist for %zu input frames",
				i, j->first_output_frame, j->first_output_frame + j->output_frames, available, j->total_input_frames);
			return -1;
		}
		if (j->output_frames && (!j->input || !j->output)) { crb_set_error("job %zu has a null buffer", i); return -1; }
	}
	for (i = 0; i < n;) {
		const ClownResamplerB200_Job *j = &jobs[i];
		crb_device_job *d = &out[n_dev++];
		size_t run = 1, k;
		uint32_t log_streams = 0;
		if (plan->kernel_kind == 0 && plan->geo.unstretched5 && j->output_frames && !no_lockstep()) {
			while (run < CRB_MAX_LOCKSTEP && i + run < n && jobs[i + run].position_fractional == j->position_fractional
			       && jobs[i + run].first_output_frame == j->first_output_frame && jobs[i + run].output_frames == j->output_frames)
				++run;
			log_streams = run >= 4 && plan->geo.lock_slot_bytes[2] ? 2 : run >= 2 && plan->geo.lock_slot_bytes[1] ? 1 : 0;
		}
		run = (size_t)1 << log_streams;
		memset(d, 0, sizeof *d);
		d->in = j->input;
		d->out = j->output;
		/* the integer position only offsets the window inside the stream's own buffer: fold it into the pointer-relative q0 of
		   stream 0 and into the input pointers of the others */
		d->q0 = ((uint64_t)j->position_integer << 16) + j->position_fractional + plan->geo.delta;
		d->first_out = j->first_output_frame;
		d->n_out = j->output_frames;
		d->in_frames = j->total_input_frames + 2 * R;
		d->increment = 0;
		d->n_more = (uint32_t)run - 1;
		for (k = 1; k < run; ++k) {
			/* stream k reads through stream 0's positions: shift its base so that its own integer position lines up */
			const ClownResamplerB200_Job *m = &jobs[i + k];
			d->in_more[k - 1] = m->input + ((ptrdiff_t)m->position_integer - (ptrdiff_t)j->position_integer) * (ptrdiff_t)plan->geo.channels;
			d->out_more[k - 1] = m->output;
			d->in_frames_more[k - 1] = m->total_input_frames + 2 * R + j->position_integer - m->position_integer;
		}
		d->tile_base = tiles;
		tiles += (j->output_frames + (plan->geo.tile_out >> log_streams) - 1) / (plan->geo.tile_out >> log_streams);
		i += run;
	}
	*n_device = n_dev;
	if (getenv("CRB200_TRACE")) fprintf(stderr, "clownresampler_b200: %zu jobs -> %zu device jobs, %llu tiles\n", n, n_dev, (unsigned long long)tiles);
	return (int64_t)tiles;
}

int ClownResamplerB200_ResampleDevice(ClownResamplerB200_Plan *plan, const ClownResamplerB200_Job *jobs,
	size_t job_count, int output_format, void *cuda_stream)
{
	crb_device_job stack_jobs[16], *dj = stack_jobs;
	int64_t tiles;
	size_t i, n_device = 0;
	int rc;
	if (!plan || (!jobs && job_count)) { crb_set_error("null argument"); return CRB200_E_ARGUMENT; }
	if (output_format < 0 || output_format > 2) { crb_set_error("unknown output format %d", output_format); return CRB200_E_ARGUMENT; }
	if (job_count == 0) return CRB200_OK;
	if (job_count > 16 && !(dj = (crb_device_job *)malloc(job_count * sizeof *dj))) { crb_set_error("out of host memory"); return CRB200_E_MEMORY; }
	tiles = convert_jobs(plan, jobs, job_count, dj, &n_device);
	if (tiles < 0) { rc = CRB200_E_ARGUMENT; goto done; }
	if (plan->kernel_kind == 0) {
		/* the vector loads of the 2-, 4-, 6- and 8-channel kernels need the frames aligned to the largest power of two dividing the
		   frame size (4, 8, 4, 16 bytes); every other channel count loads sample by sample (2 bytes) */
		const unsigned chn = plan->geo.channels;
		const size_t align = chn == 8 ? 16 : chn == 4 ? 8 : (chn == 2 || chn == 6) ? 4 : 2;
		for (i = 0; i < job_count; ++i)
			if (jobs[i].output_frames && ((uintptr_t)jobs[i].input % align) != 0) {
				crb_set_error("job %zu: input pointer must be aligned to %zu bytes", i, align);
				rc = CRB200_E_ARGUMENT; goto done;
			}
	}
	{
		const int prev = crb_dev_push(plan->device);
		rc = crb_dev_launch(plan, dj, n_device, (uint64_t)tiles, output_format, cuda_stream);
		crb_dev_pop(prev);
	}
done:
	if (dj != stack_jobs) free(dj);
	return rc;
}

/* helpers for callers that do not link the CUDA runtime themselves; allocations go to the default device unless one is named */
void *ClownResamplerB200_DeviceAllocOn(int device, size_t bytes)
{
	void *p;
	int prev;
	if ((device = crb_dev_init(device, 0)) < 0) return NULL;
	prev = crb_dev_push(device);
	p = crb_dev_alloc(bytes);
	crb_dev_pop(prev);
	return p;
}
void *ClownResamplerB200_DeviceAlloc(size_t bytes) { return ClownResamplerB200_DeviceAllocOn(-1, bytes); }
void ClownResamplerB200_DeviceFree(void *p) { crb_dev_free(p); }
void *ClownResamplerB200_PinnedAlloc(size_t bytes) { return default_device() < 0 ? NULL : crb_dev_pinned_alloc(bytes); }
void ClownResamplerB200_PinnedFree(void *p) { crb_dev_pinned_free(p); }

/* runs `body` with the device that owns `device_pointer` current */
#define ON_DEVICE_OF(device_pointer, body) do { \
		const int dev_ = crb_dev_of_pointer(device_pointer); \
		const int prev_ = dev_ >= 0 ? crb_dev_push(dev_) : -1; \
		body; \
		crb_dev_pop(prev_); \
	} while (0)

int ClownResamplerB200_CopyToDevice(void *d, const void *h, size_t bytes)
{
	int rc;
	ON_DEVICE_OF(d, { rc = crb_dev_h2d(d, h, bytes, NULL); if (!rc) rc = crb_dev_sync(NULL); });
	return rc;
}
int ClownResamplerB200_CopyToHost(void *h, const void *d, size_t bytes)
{
	int rc;
	ON_DEVICE_OF(d, { rc = crb_dev_d2h(h, d, bytes, NULL); if (!rc) rc = crb_dev_sync(NULL); });
	return rc;
}
/* waits for `cuda_stream`; a NULL stream means the default stream of `device` (-1: the default device) */
int ClownResamplerB200_SynchronizeOn(int device, void *stream)
{
	int rc, prev;
	if ((device = crb_dev_init(device, 0)) < 0) return CRB200_E_NO_DEVICE;
	prev = crb_dev_push(device);
	rc = crb_dev_sync(stream);
	crb_dev_pop(prev);
	return rc;
}
int ClownResamplerB200_Synchronize(void *stream) { return ClownResamplerB200_SynchronizeOn(-1, stream); }

/* =========================================================================================
 * format steps either side of the path: planar (one buffer per channel) device I/O
 * ========================================================================================= */
/* A stream whose channels live in separate planes is `channels` mono streams that walk through the same positions: exactly
   the lockstep jobs of the mono kernel (one phase-row fetch serves four channels).  `plan` must be a MONO plan of the same
   rates; every plane follows the padding contract on its own (H:725-733). */
int ClownResamplerB200_ResamplePlanarDevice(ClownResamplerB200_Plan *plan, const ClownResamplerB200_PlanarJob *jobs,
	size_t job_count, int output_format, void *cuda_stream)
{
	ClownResamplerB200_Job *flat;
	size_t i, c, n = 0, at = 0;
	int rc;
	if (!plan || (!jobs && job_count)) { crb_set_error("null argument"); return CRB200_E_ARGUMENT; }
	if (plan->geo.channels != 1) { crb_set_error("ResamplePlanarDevice needs a mono plan (every plane is one channel); this plan has %u channels", plan->geo.channels); return CRB200_E_ARGUMENT; }
	for (i = 0; i < job_count; ++i) {
		if (jobs[i].channels == 0 || jobs[i].channels > CLOWNRESAMPLER_MAXIMUM_CHANNELS || !jobs[i].input_planes || !jobs[i].output_planes) { crb_set_error("planar job %zu: bad planes", i); return CRB200_E_ARGUMENT; }
		n += jobs[i].channels;
	}
	if (n == 0) return CRB200_OK;
	if (!(flat = (ClownResamplerB200_Job *)calloc(n, sizeof *flat))) { crb_set_error("out of host memory"); return CRB200_E_MEMORY; }
	for (i = 0; i < job_count; ++i)
		for (c = 0; c < jobs[i].channels; ++c, ++at) {
			flat[at].input = jobs[i].input_planes[c];
			flat[at].output = jobs[i].output_planes[c];
			flat[at].total_input_frames = jobs[i].total_input_frames;
			flat[at].position_integer = jobs[i].position_integer;
			flat[at].position_fractional = jobs[i].position_fractional;
			flat[at].first_output_frame = jobs[i].first_output_frame;
			flat[at].output_frames = jobs[i].output_frames;
		}
	rc = ClownResamplerB200_ResampleDevice(plan, flat, n, output_format, cuda_stream);
	free(flat);
	return rc;
}

static int interleave_common(void *const *planes, void *interleaved, size_t frames, unsigned channels, int word_bytes, int to_planes, void *cuda_stream)
{
	int rc;
	unsigned c;
	if (!planes || !interleaved || channels == 0 || channels > CLOWNRESAMPLER_MAXIMUM_CHANNELS || (word_bytes != 2 && word_bytes != 4)) { crb_set_error("bad argument"); return CRB200_E_ARGUMENT; }
	for (c = 0; c < channels; ++c) if (!planes[c]) { crb_set_error("plane %u is null", c); return CRB200_E_ARGUMENT; }
	if (default_device() < 0) return CRB200_E_NO_DEVICE;
	ON_DEVICE_OF(interleaved, rc = crb_dev_interleave(planes, interleaved, frames, channels, word_bytes, to_planes, cuda_stream));
	return rc;
}

int ClownResamplerB200_DeinterleaveDevice(const void *interleaved, void *const *planes, size_t frames, unsigned channels, int word_bytes, void *cuda_stream)
{
	return interleave_common(planes, (void *)interleaved, frames, channels, word_bytes, 1, cuda_stream);
}

int ClownResamplerB200_InterleaveDevice(const void *const *planes, void *interleaved, size_t frames, unsigned channels, int word_bytes, void *cuda_stream)
{
	return interleave_common((void *const *)planes, interleaved, frames, channels, word_bytes, 0, cuda_stream);
}

int ClownResamplerB200_FillNoiseDevice(cc_s16l *device_dst, unsigned seed, unsigned stream, size_t first_frame, size_t n_frames, unsigned channels, void *cuda_stream)
{
	int rc;
	if (default_device() < 0) return CRB200_E_NO_DEVICE;
	ON_DEVICE_OF(device_dst, rc = crb_dev_fill_noise(device_dst, seed, stream, first_frame, n_frames, channels, cuda_stream));
	return rc;
}

int ClownResamplerB200_ChecksumDevice(const void *device_src, size_t words, int word_bytes, unsigned long *host_result, void *cuda_stream)
{
	unsigned long long v = 0;
	int rc;
	if (word_bytes != 2 && word_bytes != 4) { crb_set_error("word_bytes must be 2 or 4"); return CRB200_E_ARGUMENT; }
	ON_DEVICE_OF(device_src, rc = crb_dev_checksum(device_src, words, word_bytes, &v, cuda_stream));
	*host_result = (unsigned long)v;
	return rc;
}

/* =========================================================================================
 * staging slots (callback path and host bulk path)
 * ========================================================================================= */
static int slot_reserve(crb_slot *s, size_t in_bytes, size_t out_bytes)
{
	if (!s->stream && !(s->stream = crb_dev_stream_create())) { crb_set_error("cannot create a CUDA stream"); return CRB200_E_CUDA; }
	if (!s->done && !(s->done = crb_dev_event_create())) { crb_set_error("cannot create a CUDA event"); return CRB200_E_CUDA; }
	if (in_bytes > s->in_cap) {
		size_t cap = s->in_cap ? s->in_cap : 65536;
		while (cap < in_bytes) cap *= 2;
		crb_dev_pinned_free(s->pin_in); crb_dev_free(s->dev_in);
		s->pin_in = crb_dev_pinned_alloc(cap); s->dev_in = crb_dev_alloc(cap + 64);
		s->in_cap = (s->pin_in && s->dev_in) ? cap : 0;
		if (!s->in_cap) return CRB200_E_MEMORY;
	}
	if (out_bytes > s->out_cap) {
		size_t cap = s->out_cap ? s->out_cap : 65536;
		while (cap < out_bytes) cap *= 2;
		crb_dev_pinned_free(s->pin_out); crb_dev_free(s->dev_out);
		s->pin_out = crb_dev_pinned_alloc(cap); s->dev_out = crb_dev_alloc(cap);
		s->out_cap = (s->pin_out && s->dev_out) ? cap : 0;
		if (!s->out_cap) return CRB200_E_MEMORY;
	}
	return 0;
}

/* Submits output frames [n0, n0 + count) of the stream described by `st` / `input` (host, padded)
   on slot s: copies the input slice the frames need, runs the kernel, copies the frames back.
   The work is asynchronous; slot_wait() completes it. */
static int slot_submit(crb_slot *s, struct ClownResamplerB200_Plan *plan, const ClownResampler_LowLevel_State *st,
	const cc_s16l *input, size_t total_input_frames, size_t n0, size_t count, int fmt, int input_is_pinned, void *pinned_output)
{
	const size_t ch = plan->geo.channels, R = plan->geo.radius_int;
	const u128 p0 = position_of(st, n0), p1 = position_of(st, n0 + count - 1);
	const size_t first_in = (size_t)(p0 >> 16);
	size_t last_in = (size_t)(p1 >> 16) + 2 * R;
	crb_device_job job;
	size_t in_bytes, out_bytes;
	int rc;
	memset(&job, 0, sizeof job);
	if (last_in > total_input_frames + 2 * R) last_in = total_input_frames + 2 * R;
	in_bytes = (last_in - first_in) * ch * sizeof(cc_s16l);
	out_bytes = count * out_frame_bytes(plan, fmt);
	if ((rc = slot_reserve(s, in_bytes, out_bytes)) != 0) return rc;
	if (input_is_pinned) {
		/* page-locked caller memory: DMA straight from it */
		if ((rc = crb_dev_h2d(s->dev_in, input + first_in * ch, in_bytes, s->stream)) != 0) return rc;
	} else {
		memcpy(s->pin_in, input + first_in * ch, in_bytes);
		if ((rc = crb_dev_h2d(s->dev_in, s->pin_in, in_bytes, s->stream)) != 0) return rc;
	}
	job.in = (const int16_t *)s->dev_in;
	job.out = s->dev_out;
	job.q0 = (uint64_t)(p0 - ((u128)first_in << 16)) + plan->geo.delta;   /* position of frame n0 relative to the slice */
	job.first_out = 0;
	job.n_out = count;
	job.in_frames = last_in - first_in;
	job.increment = st->increment;       /* the plan's tiles may be sized for a larger step */
	job.tile_base = 0;
	if ((rc = crb_dev_launch(plan, &job, 1, (count + plan->geo.tile_out - 1) / plan->geo.tile_out, fmt, s->stream)) != 0) return rc;
	if ((rc = crb_dev_d2h(pinned_output ? pinned_output : s->pin_out, s->dev_out, out_bytes, s->stream)) != 0) return rc;
	return crb_dev_event_record(s->done, s->stream);
}

static int slot_wait(crb_slot *s) { return crb_dev_event_sync(s->done); }

/* =========================================================================================
 * the drop-in frame loop (H:1058-1092)
 * ========================================================================================= */
#define FIRST_CHUNK 4096u
#define MAX_CHUNK (1u << 18)

/* Frames the GPU computed ahead of a callback that then said stop (a mixer taking one tick's worth per call,
   H:746-748) are kept: the next call on the same stream is served from them without a GPU round trip -- but only
   after comparing, byte for byte, the input those frames were computed from with the input the caller presents now
   (the snapshot below), so a hit returns exactly what a fresh computation would.  One entry per state address
   (a hint only; the comparison decides), all under G.lock. */
#define MEMO_ENTRIES 4096                  /* open-addressed by state address, 8 probes; buffers are allocated lazily */
#define MEMO_PROBES 8
#define MEMO_TOTAL_BYTES ((size_t)512 << 20) /* of kept frames and snapshots over all entries */
#define MEMO_MAX_BYTES (1u << 18)          /* of cached frames per entry */
typedef struct crb_memo {
	const void *owner;                      /* state address this entry belongs to (probe key) */
	const void *state, *pre;                /* state: set while the entry holds frames */
	unsigned table;                         /* table_id() of the table the kept frames were computed with */
	ClownResampler_LowestLevel_Configuration cfg;
	cc_u8f channels;
	cc_u32f increment;
	cc_s16l *input; size_t input_frames, input_cap;   /* snapshot; frame 0 = window base of cached frame 0 */
	int32_t *frames; size_t n_frames, frames_cap;      /* cached output frames (s32), in bytes for the caps */
	uint64_t q0;                            /* 16.16 position of cached frame 0 relative to the snapshot */
	size_t next;                            /* first cached frame not delivered yet */
	int busy;                               /* a call on this state is running (only that call touches the entry) */
} crb_memo;
static crb_memo g_memo[MEMO_ENTRIES];
static size_t g_memo_bytes;

/* the entry of this state: its own if it has one, else a free one, else the first idle probe (evicted); marked busy until
   memo_done().  NULL when every candidate is in use by running calls (the call then simply keeps no frames).  Takes G.lock. */
static crb_memo *memo_for(const void *state)
{
	const size_t h = (size_t)(((uint64_t)(uintptr_t)state * 0x9E3779B97F4A7C15ull) >> 40);
	crb_memo *found = NULL, *spare = NULL, *idle = NULL;
	size_t i;
	pthread_mutex_lock(&G.lock);
	for (i = 0; i < MEMO_PROBES && !found; ++i) {
		crb_memo *m = &g_memo[(h + i) % MEMO_ENTRIES];
		if (m->owner == state) found = m;
		else if (!m->busy) {
			if (!spare && (!m->owner || m->n_frames == 0)) spare = m;
			if (!idle) idle = m;
		}
	}
	if (found && found->busy) found = NULL, spare = NULL, idle = NULL;      /* two calls on one state at once: caller's bug; keep nothing */
	else if (!found && (found = spare ? spare : idle) != NULL) {
		found->owner = state;
		found->state = NULL;
		found->n_frames = 0;
	}
	if (found) found->busy = 1;
	pthread_mutex_unlock(&G.lock);
	return found;
}

static void memo_done(crb_memo *m)
{
	if (!m) return;
	pthread_mutex_lock(&G.lock);
	m->busy = 0;
	pthread_mutex_unlock(&G.lock);
}

static unsigned long g_dropin_launches, g_memo_calls;   /* diagnostics: ClownResamplerB200_GetCounters */

void ClownResamplerB200_GetCounters(unsigned long *dropin_kernel_launches, unsigned long *calls_served_from_kept_frames)
{
	if (dropin_kernel_launches) *dropin_kernel_launches = __atomic_load_n(&g_dropin_launches, __ATOMIC_RELAXED);
	if (calls_served_from_kept_frames) *calls_served_from_kept_frames = __atomic_load_n(&g_memo_calls, __ATOMIC_RELAXED);
}

static void memo_release_all(void)
{
	int i;
	for (i = 0; i < MEMO_ENTRIES; ++i) { free(g_memo[i].input); free(g_memo[i].frames); memset(&g_memo[i], 0, sizeof g_memo[i]); }
	g_memo_bytes = 0;
}

/* reserves (or returns, when negative) bytes of the kept-frame budget */
static int memo_budget(long bytes)
{
	int ok = 1;
	pthread_mutex_lock(&G.lock);
	if (bytes > 0 && g_memo_bytes + (size_t)bytes > MEMO_TOTAL_BYTES) ok = 0;
	else g_memo_bytes = (size_t)((long)g_memo_bytes + bytes);
	pthread_mutex_unlock(&G.lock);
	return ok;
}

/* appends `n` frames (channels s32 each) as cached frames [at, at + n); frames must arrive in order */
static void memo_append(crb_memo *m, size_t at, const int32_t *frames, size_t n, size_t ch)
{
	const size_t bytes = (at + n) * ch * sizeof(int32_t);
	if (at != m->n_frames || bytes > MEMO_MAX_BYTES) return;
	if (bytes > m->frames_cap) {
		size_t cap = m->frames_cap ? m->frames_cap : 65536;
		int32_t *grown;
		while (cap < bytes) cap *= 2;
		if (!memo_budget(cap - m->frames_cap)) return;
		if (!(grown = (int32_t *)realloc(m->frames, cap))) { memo_budget(-(long)(cap - m->frames_cap)); return; }
		m->frames = grown; m->frames_cap = cap;
	}
	memcpy(m->frames + at * ch, frames, n * ch * sizeof(int32_t));
	m->n_frames = at + n;
}

cc_bool ClownResampler_LowLevel_Resample(ClownResampler_LowLevel_State *resampler,
	const ClownResampler_Precomputed *precomputed, const cc_s16l *input_buffer, size_t *total_input_frames,
	ClownResampler_OutputCallback output_callback, const void *user_data)
{
	const size_t total = *total_input_frames;
	const size_t n_total = ClownResamplerB200_CountOutputFrames(resampler, total);
	const cc_u8f ch = resampler->channels;
	const size_t R = resampler->lowest_level.integer_stretched_kernel_radius;
	struct ClownResamplerB200_Plan *plan = NULL;
	crb_lane *lane = NULL;
	crb_memo *memo;
	/* first speculative chunk: as many frames as half an entry of the kept-frame table holds, 2048..16384 (a whole
	   HighLevel refill of mono or stereo input in one launch) */
	size_t delivered = 0, submitted = 0, pending_n[SLOTS], pending_k0[SLOTS], memo_base;
	size_t chunk = MEMO_MAX_BYTES / 2 / (ch * sizeof(int32_t)) > 4 * FIRST_CHUNK ? 4 * FIRST_CHUNK
		: MEMO_MAX_BYTES / 2 / (ch * sizeof(int32_t)) < FIRST_CHUNK / 2 ? FIRST_CHUNK / 2 : MEMO_MAX_BYTES / 2 / (ch * sizeof(int32_t));
	unsigned head = 0, tail = 0; /* slots [tail, head) are in flight */
	unsigned table;
	int stopped = 0, rc = 0, device, prev_device = -1;

	if (n_total == 0) {          /* H:1063-1067 with no frame emitted */
		ClownResamplerB200_AdvanceState(resampler, total_input_frames, 0, 0);
		return cc_true;
	}
	if (ch == 0 || ch > CLOWNRESAMPLER_MAXIMUM_CHANNELS) {
		crb_set_error("channels must be 1..%d (got %u); the reference's accumulator array has %d slots (H:1071)", CLOWNRESAMPLER_MAXIMUM_CHANNELS, ch, CLOWNRESAMPLER_MAXIMUM_CHANNELS);
		report("ClownResampler_LowLevel_Resample produced no frames");
		return cc_false;
	}
	table = table_id(precomputed);
	memo = table ? memo_for(resampler) : NULL;      /* no lock is held from here on: the callbacks below may call back into the library */

	/* 1. frames computed ahead by the previous call on this stream, if the input they came from is still what the
	      caller presents */
	if (memo && memo->state == resampler && memo->next < memo->n_frames && memo->pre == precomputed && memo->channels == ch
	    && memo->increment == resampler->increment && memcmp(&memo->cfg, &resampler->lowest_level, sizeof memo->cfg) == 0
	    && table && memo->table == table) {
		const uint64_t qk = memo->q0 + (uint64_t)memo->next * resampler->increment;
		size_t m = memo->n_frames - memo->next;
		if (m > n_total) m = n_total;
		if ((qk & 0xFFFF) == resampler->position_fractional) {
			const size_t snap_first = (size_t)(qk >> 16);
			const size_t count = (size_t)((qk + (uint64_t)(m - 1) * resampler->increment) >> 16) - snap_first + 2 * R;
			if (snap_first + count <= memo->input_frames && resampler->position_integer + count <= total + 2 * R
			    && memcmp(input_buffer + resampler->position_integer * ch, memo->input + snap_first * ch, count * ch * sizeof(cc_s16l)) == 0) {
				const int32_t *frames = memo->frames + memo->next * ch;
				size_t k;
				for (k = 0; k < m && !stopped; ++k) {
					cc_s32f frame[CLOWNRESAMPLER_MAXIMUM_CHANNELS];
					cc_u8f c;
					for (c = 0; c < ch; ++c) frame[c] = frames[k * ch + c];
					++delivered;
					if (!output_callback((void *)user_data, frame, ch)) stopped = 1;
				}
				memo->next += delivered;
				__atomic_add_fetch(&g_memo_calls, 1, __ATOMIC_RELAXED);
				if (stopped || delivered == n_total) {
					memo_done(memo);
					ClownResamplerB200_AdvanceState(resampler, total_input_frames, delivered, stopped);
					return stopped ? cc_false : cc_true;
				}
			}
		}
	}
	/* 2. the GPU computes frames [delivered, n_total) in growing chunks, ahead of the callbacks */
	if (memo) { memo->n_frames = 0; memo->next = 0; memo->state = NULL; }
	memo_base = submitted = delivered;
	if ((device = default_device()) < 0) { rc = CRB200_E_NO_DEVICE; goto fail; }
	if (!(plan = plan_cached(precomputed, resampler, device, table))) { rc = CRB200_E_CONFIG; goto fail; }
	lane = lane_acquire(device);

	while (delivered < n_total && !stopped) {
		/* keep the pipeline full: the next chunk computes while this one is delivered */
		prev_device = crb_dev_push(device);
		while (submitted < n_total && head - tail < SLOTS) {
			const size_t n = n_total - submitted < chunk ? n_total - submitted : chunk;
			if ((rc = slot_submit(&lane->slots[head % SLOTS], plan, resampler, input_buffer, total, submitted, n, CRB200_OUT_S32, 0, NULL)) != 0) break;
			__atomic_add_fetch(&g_dropin_launches, 1, __ATOMIC_RELAXED);
			pending_n[head % SLOTS] = n;
			pending_k0[head % SLOTS] = submitted;
			submitted += n;
			++head;
			if (chunk < MAX_CHUNK) chunk *= 2;
		}
		if (rc == 0 && head != tail) rc = slot_wait(&lane->slots[tail % SLOTS]);
		crb_dev_pop(prev_device);        /* the callbacks run with the caller's own device current */
		if (rc != 0) goto fail;
		{
			crb_slot *s = &lane->slots[tail % SLOTS];
			const int32_t *frames = (const int32_t *)s->pin_out;
			const size_t n = pending_n[tail % SLOTS];
			size_t k;
			for (k = 0; k < n; ++k) {
				cc_s32f frame[CLOWNRESAMPLER_MAXIMUM_CHANNELS];
				cc_u8f c;
				for (c = 0; c < ch; ++c) frame[c] = frames[k * ch + c];
				++delivered;
				if (!output_callback((void *)user_data, frame, ch)) { stopped = 1; break; }
			}
			if (stopped && memo) {
				/* keep the whole chunk the callback stopped in (`next` will skip what was delivered) and, below, the
				   chunks already computed behind it */
				memo_base = pending_k0[tail % SLOTS];
				memo_append(memo, 0, frames, n, ch);
			}
			++tail;
		}
	}
	/* chunks computed ahead of a callback that stopped: kept for the next call on this stream */
	prev_device = crb_dev_push(device);
	while (tail < head) {
		crb_slot *s = &lane->slots[tail % SLOTS];
		if (slot_wait(s) == 0 && stopped && memo)
			memo_append(memo, pending_k0[tail % SLOTS] - memo_base, (const int32_t *)s->pin_out, pending_n[tail % SLOTS], ch);
		++tail;
	}
	crb_dev_pop(prev_device);
	if (memo && stopped && memo->n_frames > delivered - memo_base) {
		/* cached frame 0 is this call's frame memo_base; snapshot the input its successors were computed from */
		const u128 p_first = position_of(resampler, memo_base), p_last = position_of(resampler, memo_base + memo->n_frames - 1);
		const size_t base_frame = (size_t)(p_first >> 16);
		size_t end_frame = (size_t)(p_last >> 16) + 2 * R;
		size_t bytes;
		if (end_frame > total + 2 * R) end_frame = total + 2 * R;
		bytes = (end_frame - base_frame) * ch * sizeof(cc_s16l);
		if (bytes > memo->input_cap && memo_budget((long)(bytes * 2 - memo->input_cap))) {
			cc_s16l *grown = (cc_s16l *)realloc(memo->input, bytes * 2);
			if (grown) { memo->input = grown; memo->input_cap = bytes * 2; }
			else memo_budget(-(long)(bytes * 2 - memo->input_cap));
		}
		if (bytes <= memo->input_cap) {
			memcpy(memo->input, input_buffer + base_frame * ch, bytes);
			memo->input_frames = end_frame - base_frame;
			memo->q0 = (uint64_t)(p_first - ((u128)base_frame << 16));
			memo->next = delivered - memo_base;
			memo->state = resampler; memo->pre = precomputed; memo->table = table;
			memo->cfg = resampler->lowest_level; memo->channels = ch; memo->increment = resampler->increment;
		} else {
			memo->n_frames = 0;
		}
	} else if (memo) {
		memo->n_frames = 0;
	}
	memo_done(memo);
	lane_release(lane);
	plan_unref(plan);
	ClownResamplerB200_AdvanceState(resampler, total_input_frames, delivered, stopped);
	return stopped ? cc_false : cc_true;

fail:
	/* No frame can be computed (no device, a configuration the reference itself cannot run, a CUDA failure): say so --
	   stderr and ClownResamplerB200_GetLastError() -- and stop WITHOUT consuming the input that produced no output:
	   the state advances over the frames already delivered, exactly as if the callback had asked to stop there
	   (H:1084-1088), and the call returns cc_false.  There is no CPU fallback. */
	if (lane) {
		prev_device = crb_dev_push(device);
		while (tail < head) { slot_wait(&lane->slots[tail % SLOTS]); ++tail; }
		crb_dev_pop(prev_device);
		lane_release(lane);
	}
	if (memo) { memo->n_frames = 0; memo->state = NULL; }
	memo_done(memo);
	if (plan) plan_unref(plan);
	report("ClownResampler_LowLevel_Resample stopped early");
	(void)rc;
	ClownResamplerB200_AdvanceState(resampler, total_input_frames, delivered, 1);
	return cc_false;
}

void ClownResampler_LowestLevel_Resample(const ClownResampler_LowestLevel_Configuration *configuration,
	const ClownResampler_Precomputed *precomputed, cc_s32f *output_frame, cc_u8f channels,
	const cc_s16l *input_buffer, size_t position_integer, cc_u32f position_fractional)
{
	/* H:986-1035 for one frame: raw accumulators from the GPU, added to the caller's
	   accumulators and normalised the way H:1020/H:1033 do (output_frame is read-modify-write). */
	ClownResampler_LowLevel_State st;
	struct ClownResamplerB200_Plan *plan = NULL;
	int rc = CRB200_E_CONFIG, device;
	cc_u8f c;
	memset(&st, 0, sizeof st);
	st.lowest_level = *configuration;
	st.channels = channels;
	st.position_integer = position_integer;
	st.position_fractional = position_fractional;
	st.increment = FX;
	if ((device = default_device()) >= 0 && (plan = plan_cached(precomputed, &st, device, table_id(precomputed))) != NULL) {
		crb_lane *lane = lane_acquire(device);
		const int prev = crb_dev_push(device);
		if ((rc = slot_submit(&lane->slots[0], plan, &st, input_buffer, position_integer + 1, 0, 1, 2, 0, NULL)) == 0 && (rc = slot_wait(&lane->slots[0])) == 0) {
			const int32_t *raw = (const int32_t *)lane->slots[0].pin_out;
			for (c = 0; c < channels; ++c)
				output_frame[c] = (output_frame[c] + raw[c]) * (cc_s32f)raw[channels] / (1 << 15);
		}
		crb_dev_pop(prev);
		lane_release(lane);
		plan_unref(plan);
	}
	if (rc != 0)
		report("ClownResampler_LowestLevel_Resample left the frame untouched");
}

/* =========================================================================================
 * host bulk path: same kernels, host pointers, copies overlapped with compute
 * ========================================================================================= */
/* Bytes (input slice + output) per pipelined chunk.  Measured on the bench workload (stereo, PCIe both ways at once): 8 M output
   frames per chunk (about 62 MB) gave 22.0 Gsamples/s end to end against 21.0 for 2 M and 18.1 for 512 K -- each chunk costs a
   fixed copy-queue turnaround.  The chunk is sized in BYTES so that wide or steeply down-sampled streams do not multiply the pinned
   and device staging (three slots per lane), and halves when an allocation fails. */
#ifndef CRB_HOST_CHUNK_BYTES
#define CRB_HOST_CHUNK_BYTES ((size_t)64 << 20)
#endif

int ClownResamplerB200_ResampleHost(ClownResamplerB200_Plan *plan, const ClownResamplerB200_Job *jobs, size_t job_count, int output_format)
{
	size_t j;
	int rc = 0, prev;
	unsigned head = 0, tail = 0;
	struct { unsigned char *dst; size_t bytes; } pend[SLOTS];
	crb_lane *lane;
	size_t chunk_bytes = CRB_HOST_CHUNK_BYTES;
	if (!plan || (!jobs && job_count)) { crb_set_error("null argument"); return CRB200_E_ARGUMENT; }
	if (output_format < 0 || output_format > 2) { crb_set_error("unknown output format %d", output_format); return CRB200_E_ARGUMENT; }
	lane = lane_acquire(plan->device);
	prev = crb_dev_push(plan->device);
	for (j = 0; j < job_count && rc == 0; ++j) {
		const ClownResamplerB200_Job *job = &jobs[j];
		ClownResampler_LowLevel_State st;
		size_t done = 0;
		const size_t fb_out = out_frame_bytes(plan, output_format);
		/* bytes one output frame moves: its own, plus its share of the input (increment / 65536 input frames) */
		const double bytes_per_frame = (double)fb_out + (double)plan->geo.increment / 65536.0 * 2.0 * plan->geo.channels;
		const int in_pinned = job->output_frames && crb_dev_is_pinned(job->input, (job->total_input_frames + 2 * plan->geo.radius_int) * plan->geo.channels * sizeof(cc_s16l));
		const int out_pinned = job->output_frames && crb_dev_is_pinned(job->output, job->output_frames * fb_out);
		memset(&st, 0, sizeof st);
		st.position_integer = job->position_integer;
		st.position_fractional = job->position_fractional;
		st.increment = plan->geo.increment;
		if (job->first_output_frame + job->output_frames > ClownResamplerB200_CountOutputFrames(&st, job->total_input_frames)) {
			crb_set_error("job %zu asks for more output frames than its input yields", j);
			rc = CRB200_E_ARGUMENT;
			break;
		}
		while (done < job->output_frames && rc == 0) {
			size_t n = (size_t)((double)chunk_bytes / bytes_per_frame);
			if (n < 4096) n = 4096;
			if (n > job->output_frames - done) n = job->output_frames - done;
			if (head - tail == SLOTS) {
				crb_slot *s = &lane->slots[tail % SLOTS];
				if ((rc = slot_wait(s)) != 0) break;
				if (pend[tail % SLOTS].dst) memcpy(pend[tail % SLOTS].dst, s->pin_out, pend[tail % SLOTS].bytes);
				++tail;
			}
			rc = slot_submit(&lane->slots[head % SLOTS], plan, &st, job->input, job->total_input_frames, job->first_output_frame + done, n, output_format,
				in_pinned, out_pinned ? (unsigned char *)job->output + done * fb_out : NULL);
			if (rc == CRB200_E_MEMORY && chunk_bytes > ((size_t)1 << 20)) {
				/* staging did not fit: work in smaller pieces */
				chunk_bytes /= 2;
				rc = 0;
				continue;
			}
			pend[head % SLOTS].dst = out_pinned ? NULL : (unsigned char *)job->output + done * fb_out;
			pend[head % SLOTS].bytes = n * fb_out;
			if (rc == 0) ++head;
			done += n;
		}
	}
	while (tail < head) {
		crb_slot *s = &lane->slots[tail % SLOTS];
		const int w = slot_wait(s);
		if (w == 0 && rc == 0 && pend[tail % SLOTS].dst) memcpy(pend[tail % SLOTS].dst, s->pin_out, pend[tail % SLOTS].bytes);
		if (w != 0 && rc == 0) rc = w;
		++tail;
	}
	crb_dev_pop(prev);
	lane_release(lane);
	return rc;
}

/* =========================================================================================
 * all GPUs of a box from one plain-C call: the path shards without any exchange step
 * ========================================================================================= */
typedef struct crb_multi_worker {
	const ClownResampler_Precomputed *pre;
	const ClownResampler_LowLevel_State *state;
	int device, output_format, rc;
	ClownResamplerB200_Job *jobs;
	size_t job_count;
	char error[512];
} crb_multi_worker;

static void *multi_worker_main(void *arg)
{
	crb_multi_worker *w = (crb_multi_worker *)arg;
	struct ClownResamplerB200_Plan *plan = NULL;
	w->rc = CRB200_OK;
	if (w->job_count == 0) return NULL;
	if (crb_dev_init(w->device, 0) < 0) w->rc = CRB200_E_NO_DEVICE;
	else if (!(plan = plan_create(w->pre, w->state, w->state->increment, w->device))) w->rc = CRB200_E_CONFIG;   /* exact increment: the bulk paths take it from the plan */
	else {
		w->rc = ClownResamplerB200_ResampleHost(plan, w->jobs, w->job_count, w->output_format);
		plan_unref(plan);
	}
	if (w->rc != CRB200_OK) snprintf(w->error, sizeof w->error, "device %d: %s", w->device, ClownResamplerB200_GetLastError());
	return NULL;
}

int ClownResamplerB200_ResampleHostMulti(const ClownResampler_Precomputed *precomputed, const ClownResampler_LowLevel_State *state,
	const int *devices, size_t device_count, const ClownResamplerB200_Job *jobs, size_t job_count, int output_format)
{
	crb_multi_worker *workers;
	pthread_t *threads;
	ClownResamplerB200_Job *parts;
	size_t d, j, n_parts, fb_out;
	int rc = CRB200_OK;
	if (!precomputed || !state || !devices || device_count == 0 || (!jobs && job_count)) { crb_set_error("null argument"); return CRB200_E_ARGUMENT; }
	if (output_format != CRB200_OUT_S32 && output_format != CRB200_OUT_S16_CLAMPED) { crb_set_error("unknown output format %d", output_format); return CRB200_E_ARGUMENT; }
	if (job_count == 0) return CRB200_OK;
	fb_out = (output_format == CRB200_OUT_S16_CLAMPED ? 2u : 4u) * state->channels;
	/* enough independent streams: deal them in contiguous blocks (sizes differ by at most one).  Fewer streams than devices:
	   cut every stream into device_count contiguous output-time segments instead -- a segment needs only its own slice of the
	   input plus the kernel-radius halo (H:725-733), which the host path uploads from the one buffer all segments share */
	n_parts = job_count >= device_count ? job_count : job_count * device_count;
	workers = (crb_multi_worker *)calloc(device_count, sizeof *workers);
	threads = (pthread_t *)calloc(device_count, sizeof *threads);
	parts = (ClownResamplerB200_Job *)calloc(n_parts, sizeof *parts);
	if (!workers || !threads || !parts) { free(workers); free(threads); free(parts); crb_set_error("out of host memory"); return CRB200_E_MEMORY; }
	if (job_count >= device_count) {
		const size_t base = job_count / device_count, extra = job_count % device_count;
		size_t at = 0;
		memcpy(parts, jobs, job_count * sizeof *parts);
		for (d = 0; d < device_count; ++d) {
			workers[d].jobs = parts + at;
			workers[d].job_count = base + (d < extra ? 1 : 0);
			at += workers[d].job_count;
		}
	} else {
		for (d = 0; d < device_count; ++d) {
			workers[d].jobs = parts + d * job_count;
			workers[d].job_count = job_count;
			for (j = 0; j < job_count; ++j) {
				ClownResamplerB200_Job *p = &workers[d].jobs[j];
				const size_t n0 = (size_t)((u128)jobs[j].output_frames * d / device_count), n1 = (size_t)((u128)jobs[j].output_frames * (d + 1) / device_count);
				*p = jobs[j];
				p->first_output_frame = jobs[j].first_output_frame + n0;
				p->output_frames = n1 - n0;
				p->output = (unsigned char *)jobs[j].output + n0 * fb_out;
			}
		}
	}
	for (d = 0; d < device_count; ++d) {
		workers[d].pre = precomputed; workers[d].state = state; workers[d].device = devices[d]; workers[d].output_format = output_format;
		if (pthread_create(&threads[d], NULL, multi_worker_main, &workers[d]) != 0) { multi_worker_main(&workers[d]); threads[d] = 0; }
	}
	for (d = 0; d < device_count; ++d) {
		if (threads[d]) pthread_join(threads[d], NULL);
		if (workers[d].rc != CRB200_OK && rc == CRB200_OK) { rc = workers[d].rc; crb_set_error("%s", workers[d].error); }
	}
	free(workers); free(threads); free(parts);
	return rc;
}

/* =========================================================================================
 * streaming wrapper (H:1101-1250): host-side buffer management only; every frame comes from
 * ClownResampler_LowLevel_Resample above.  Buffer layout, as in the reference:
 *   [ R carried frames | R look-ahead frames | freshly pulled frames ... ]   R = radius at Init
 * ========================================================================================= */
#define HL_SAMPLES (sizeof(((ClownResampler_HighLevel_State *)0)->input_buffer) / sizeof(cc_s16l))

cc_bool ClownResampler_HighLevel_Init(ClownResampler_HighLevel_State *resampler, cc_u8f channels,
	cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate)
{
	size_t R;
	if (channels > CLOWNRESAMPLER_MAXIMUM_CHANNELS)                                                                          /* H:1103 */
		return cc_false;
	if (!ClownResampler_LowLevel_Init(&resampler->low_level, channels, input_sample_rate, output_sample_rate, low_pass_filter_sample_rate))
		return cc_false;
	R = resampler->low_level.lowest_level.integer_stretched_kernel_radius;
	if (channels == 0 || 2 * R * channels >= HL_SAMPLES) {
		/* the reference does not check this at Init (only in Adjust, H:1202) and overruns its buffer; refuse instead */
		crb_set_error("kernel radius %zu x %u channels does not fit the %zu-sample streaming buffer", R, channels, (size_t)HL_SAMPLES);
		return cc_false;
	}
	resampler->maximum_integer_stretched_kernel_radius = R;                                                                  /* H:1109 */
	resampler->leading_padding_frames_needed = R;
	resampler->trailing_padding_frames_remaining = R;
	memset(resampler->input_buffer, 0, R * channels * sizeof(cc_s16l));                                                       /* H:1112 */
	resampler->input_buffer_start = resampler->input_buffer_end = resampler->input_buffer + R * channels;                    /* H:1115 */
	return cc_true;
}

cc_bool ClownResampler_HighLevel_Resample(ClownResampler_HighLevel_State *resampler,
	const ClownResampler_Precomputed *precomputed, ClownResampler_InputCallback input_callback,
	ClownResampler_OutputCallback output_callback, const void *user_data)
{
	const size_t ch = resampler->low_level.channels;
	const size_t halo = resampler->maximum_integer_stretched_kernel_radius * ch;   /* samples in one dead zone */
	cc_s16l *const buf = resampler->input_buffer;

	/* the first R frames of the stream go straight into the look-ahead zone (H:1127-1136) */
	while (resampler->leading_padding_frames_needed != 0) {
		const size_t want = resampler->leading_padding_frames_needed;
		const size_t got = input_callback((void *)user_data, buf + 2 * halo - want * ch, want);
		if (got == 0)
			return cc_true;
		resampler->leading_padding_frames_needed -= got;
	}
	for (;;) {
		size_t frames;
		if (resampler->input_buffer_start == resampler->input_buffer_end) {
			/* refill: carry the last 2R frames to the front, append new frames (H:1141-1158) */
			size_t got;
			memmove(buf, resampler->input_buffer_end - halo, 2 * halo * sizeof(cc_s16l));
			resampler->input_buffer_start = buf + halo;
			got = input_callback((void *)user_data, buf + 2 * halo, (HL_SAMPLES - 2 * halo) / ch);
			resampler->input_buffer_end = resampler->input_buffer_start + got * ch;
			if (got == 0)
				return cc_true;
		}
		frames = (size_t)(resampler->input_buffer_end - resampler->input_buffer_start) / ch;                              /* H:1167 */
		{
			const size_t pad = resampler->low_level.lowest_level.integer_stretched_kernel_radius * ch;
			const cc_bool ran_out = ClownResampler_LowLevel_Resample(&resampler->low_level, precomputed,
				resampler->input_buffer_start - pad, &frames, output_callback, user_data);
			resampler->input_buffer_start = resampler->input_buffer_end - frames * ch;                                     /* H:1171 */
			if (!ran_out)
				return cc_false;
		}
	}
}

cc_bool ClownResampler_HighLevel_Adjust(ClownResampler_HighLevel_State *resampler,
	cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate)
{
	/* H:1183-1209 */
	const ClownResampler_LowLevel_State saved = resampler->low_level;
	if (ClownResampler_LowLevel_Adjust(&resampler->low_level, input_sample_rate, output_sample_rate, low_pass_filter_sample_rate)) {
		const size_t R = resampler->low_level.lowest_level.integer_stretched_kernel_radius;
		if (R <= resampler->maximum_integer_stretched_kernel_radius && R * 2 < HL_SAMPLES / resampler->low_level.channels)
			return cc_true;
	}
	resampler->low_level = saved;
	return cc_false;
}

typedef struct crb_flush {
	ClownResampler_HighLevel_State *resampler;
	ClownResampler_OutputCallback output_callback;
	void *user_data;
} crb_flush;

static size_t crb_flush_input(void *user, cc_s16l *buffer, size_t total_frames)
{
	/* H:1223-1233: the stream ends with R frames of silence */
	crb_flush *f = (crb_flush *)user;
	size_t n = f->resampler->trailing_padding_frames_remaining;
	if (n > total_frames) n = total_frames;
	memset(buffer, 0, n * f->resampler->low_level.channels * sizeof(cc_s16l));
	f->resampler->trailing_padding_frames_remaining -= n;
	return n;
}

static cc_bool crb_flush_output(void *user, const cc_s32f *frame, cc_u8f total_samples)
{
	crb_flush *f = (crb_flush *)user;
	return f->output_callback(f->user_data, frame, total_samples);
}

cc_bool ClownResampler_HighLevel_ResampleEnd(ClownResampler_HighLevel_State *resampler,
	const ClownResampler_Precomputed *precomputed, ClownResampler_OutputCallback output_callback, const void *user_data)
{
	crb_flush f;
	f.resampler = resampler;
	f.output_callback = output_callback;
	f.user_data = (void *)user_data;
	return ClownResampler_HighLevel_Resample(resampler, precomputed, crb_flush_input, crb_flush_output, &f);
}
