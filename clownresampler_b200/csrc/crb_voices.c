/*
 * crb_voices.c -- batched streaming front end: many HighLevel-style voices advanced per kernel launch.
 *
 * The reference's wrapper (H:1101-1250, H = /root/reference/clownresampler.h) feeds its low-level loop from a
 * 4096-sample window buffer: [R carried frames | R look-ahead frames | new frames], i.e. the stream is the
 * low-level resampler run over  R zero frames + input + R zero frames,  in whatever chunks the callbacks
 * deliver (the reference's own test suite pins "wrapper == one shot": both harnesses share their goldens,
 * tests/CMakeLists.txt:25-47).  So a voice is fully described by that padded stream and the number of output
 * frames already emitted; positions come from the closed form, and one tick of ALL voices is one batch of
 * independent jobs for the tiled kernel.  Host code only moves bytes; every frame is computed on the GPU.
 *
 * R above is the radius the voice was created with (maximum_integer_stretched_kernel_radius, H:1109).  A voice whose
 * ratio is adjusted mid-stream (ClownResampler_HighLevel_Adjust, H:1183-1209) may run a narrower kernel of radius
 * r <= R afterwards: the low-level loop is then handed the buffer r frames before the stream position (H:1168), i.e.
 * its window base moves R - r frames into the padded stream.  Voices are grouped by kernel geometry per tick: one
 * launch per geometry in use (one plan each, built on first use), all behind the same upload and before the same download.
 */
#include "../../include/clownresampler_b200.h"
#include "crb_internal.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef unsigned __int128 u128;

#define VOICE_PLANS 16           /* distinct kernel geometries alive in one batch */
#define HL_SAMPLES 0x1000u       /* the reference wrapper's buffer, H:653: bounds the radius an Adjust may ask for (H:1202) */

/* CRB200_TRACE=1: per-phase wall time of the ticks, printed when the batch is destroyed */
static double crb_now(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

typedef struct crb_voice {
	cc_s16l *data;          /* allocation; the live frames start `off` frames in */
	size_t off;             /* frames at the front of the allocation that no later window can read (dropped lazily) */
	size_t base;            /* padded-stream index of the first live frame: R zeros, pushed input, (R zeros once ended) */
	size_t frames;          /* frames stored in data */
	size_t capacity;        /* frames allocated */
	size_t pushed;          /* input frames pushed so far */
	size_t pos_int;         /* position of the next output frame in the padded stream (H:645-646) */
	cc_u32f pos_frac;
	cc_u32f increment;      /* this voice's 16.16 step (H:647); changed by VoiceBatchAdjust */
	unsigned radius;        /* this voice's current integer_stretched_kernel_radius (<= the batch's) */
	int plan;               /* index into the batch's plan table */
	int ended;
} crb_voice;

typedef struct crb_voice_plan {
	ClownResamplerB200_Plan *plan;
	ClownResampler_LowestLevel_Configuration cfg;
	cc_u32f increment;      /* the increment the plan's tiles are sized for: >= every voice's that uses it */
} crb_voice_plan;

/* what one active voice contributes to the tick in flight */
typedef struct crb_voice_slice {
	size_t first, last;     /* padded-stream frames [first, last) uploaded for it */
	const int16_t *dev_in;
	void *dev_out;
	uint64_t q0;
} crb_voice_slice;

struct ClownResamplerB200_VoiceBatch {
	int device;                             /* the device the plans, the stream and the staging buffers live on */
	ClownResampler_Precomputed *table;      /* copy of the caller's table, for plans built later */
	ClownResampler_LowLevel_State init;     /* configuration the batch was created with */
	size_t voices, channels, radius;        /* radius: R of the creation rates = every voice's maximum radius */
	crb_voice *voice;
	crb_voice_plan plans[VOICE_PLANS];
	int n_plans;
	/* per-tick staging */
	void *stream;
	cc_s16l *pin_in; void *dev_in; size_t in_cap;
	unsigned char *pin_out; void *dev_out; size_t out_cap;
	crb_voice_slice *slice;
	/* the tick in flight (TickBegin .. TickEnd) */
	int in_flight, direct, fmt;
	void *out; size_t out_stride; size_t *produced;
	int trace; double t_plan, t_gather, t_device, t_scatter; size_t ticks;
};

#define LIVE(b, v) ((v)->data + (v)->off * (b)->channels)

static int voice_reserve(ClownResamplerB200_VoiceBatch *b, crb_voice *v, size_t extra)
{
	if (v->off + v->frames + extra > v->capacity && v->off) {
		/* consumed frames are dropped by moving an offset; the live ones move to the front only when room runs out */
		memmove(v->data, LIVE(b, v), v->frames * b->channels * sizeof(cc_s16l));
		v->off = 0;
	}
	if (v->frames + extra > v->capacity) {
		size_t cap = v->capacity ? v->capacity : 4096;
		cc_s16l *p;
		while (cap < v->frames + extra) cap *= 2;
		p = (cc_s16l *)realloc(v->data, cap * b->channels * sizeof(cc_s16l));
		if (!p) { crb_set_error("out of host memory"); return CRB200_E_MEMORY; }
		v->data = p;
		v->capacity = cap;
	}
	return 0;
}

/* the plan for configuration `cfg` with tiles sized for at least `increment`: an existing one, else a new one */
static int plan_for(ClownResamplerB200_VoiceBatch *b, const ClownResampler_LowestLevel_Configuration *cfg, cc_u32f increment)
{
	ClownResampler_LowLevel_State st;
	ClownResamplerB200_Plan *plan;
	int i, slot = -1;
	for (i = 0; i < b->n_plans; ++i)
		if (memcmp(&b->plans[i].cfg, cfg, sizeof *cfg) == 0) {
			if (b->plans[i].increment >= increment) return i;
			slot = i;            /* same geometry, tiles too small: rebuild in place for the larger step */
			break;
		}
	if (slot < 0 && b->n_plans == VOICE_PLANS) { crb_set_error("more than %d kernel geometries in one voice batch", VOICE_PLANS); return -1; }
	st = b->init;
	st.lowest_level = *cfg;
	/* up-sampling voices share tiles sized for increment 1.0; others round up a little so that small bends do not rebuild */
	st.increment = increment <= CRB_FX_ONE ? CRB_FX_ONE : increment + increment / 8;
	plan = ClownResamplerB200_PlanCreateOnDevice(b->table, &st, b->device);
	if (!plan) return -1;
	if (slot < 0) slot = b->n_plans++;
	else ClownResamplerB200_PlanDestroy(b->plans[slot].plan);
	b->plans[slot].plan = plan;
	b->plans[slot].cfg = *cfg;
	b->plans[slot].increment = st.increment;
	return slot;
}

ClownResamplerB200_VoiceBatch *ClownResamplerB200_VoiceBatchCreate(const ClownResampler_Precomputed *precomputed,
	size_t voices, cc_u8f channels, cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate)
{
	ClownResamplerB200_VoiceBatch *b;
	size_t i;
	int prev, first_plan = -1;
	if (!precomputed || voices == 0) { crb_set_error("bad argument"); return NULL; }
	b = (ClownResamplerB200_VoiceBatch *)calloc(1, sizeof *b);
	if (!b) { crb_set_error("out of host memory"); return NULL; }
	if (channels == 0 || channels > CLOWNRESAMPLER_MAXIMUM_CHANNELS                                                /* H:1103 */
	    || !ClownResampler_LowLevel_Init(&b->init, channels, input_sample_rate, output_sample_rate, low_pass_filter_sample_rate)) {
		crb_set_error("configuration rejected (channels %u, rates %lu -> %lu, low-pass %lu)", channels, input_sample_rate, output_sample_rate, low_pass_filter_sample_rate);
		free(b);
		return NULL;
	}
	b->device = crb_dev_init(-1, 0);
	b->table = (ClownResampler_Precomputed *)malloc(sizeof *b->table);
	if (b->table) *b->table = *precomputed;
	if (b->device >= 0 && b->table) first_plan = plan_for(b, &b->init.lowest_level, b->init.increment);
	b->voice = (crb_voice *)calloc(voices, sizeof *b->voice);
	b->slice = (crb_voice_slice *)malloc(voices * sizeof *b->slice);
	prev = first_plan >= 0 ? crb_dev_push(b->device) : -1;
	b->stream = first_plan >= 0 ? crb_dev_stream_create() : NULL;
	crb_dev_pop(prev);
	if (first_plan < 0 || !b->table || !b->voice || !b->slice || !b->stream) {
		if (first_plan >= 0 && (!b->table || !b->voice || !b->slice)) crb_set_error("out of host memory");
		ClownResamplerB200_VoiceBatchDestroy(b);
		return NULL;
	}
	b->trace = getenv("CRB200_TRACE") != NULL;
	b->voices = voices;
	b->channels = channels;
	b->radius = b->init.lowest_level.integer_stretched_kernel_radius;
	for (i = 0; i < voices; ++i) {
		/* the R frames of silence before the stream (H:1112) */
		crb_voice *v = &b->voice[i];
		if (voice_reserve(b, v, b->radius) != 0) { ClownResamplerB200_VoiceBatchDestroy(b); return NULL; }
		memset(LIVE(b, v), 0, b->radius * channels * sizeof(cc_s16l));
		v->frames = b->radius;
		v->increment = b->init.increment;
		v->radius = (unsigned)b->radius;
		v->plan = first_plan;
	}
	return b;
}

void ClownResamplerB200_VoiceBatchDestroy(ClownResamplerB200_VoiceBatch *b)
{
	size_t i;
	int k;
	if (!b) return;
	if (b->trace && b->ticks)
		fprintf(stderr, "clownresampler_b200 VoiceBatch: %zu ticks; per tick: plan %.1f us, gather + submit %.1f us, waiting for upload + kernel + download %.1f us, scatter %.1f us\n",
			b->ticks, 1e6 * b->t_plan / b->ticks, 1e6 * b->t_gather / b->ticks, 1e6 * b->t_device / b->ticks, 1e6 * b->t_scatter / b->ticks);
	if (b->voice) for (i = 0; i < b->voices; ++i) free(b->voice[i].data);
	free(b->voice); free(b->slice); free(b->table);
	if (b->device >= 0) {
		const int prev = crb_dev_push(b->device);
		if (b->in_flight && b->stream) crb_dev_sync(b->stream);
		crb_dev_pinned_free(b->pin_in); crb_dev_free(b->dev_in);
		crb_dev_pinned_free(b->pin_out); crb_dev_free(b->dev_out);
		crb_dev_stream_destroy(b->stream);
		crb_dev_pop(prev);
	}
	for (k = 0; k < b->n_plans; ++k) ClownResamplerB200_PlanDestroy(b->plans[k].plan);
	free(b);
}

int ClownResamplerB200_VoiceBatchPush(ClownResamplerB200_VoiceBatch *b, size_t voice, const cc_s16l *input, size_t frames)
{
	crb_voice *v;
	int rc;
	if (!b || voice >= b->voices || (!input && frames)) { crb_set_error("bad argument"); return CRB200_E_ARGUMENT; }
	v = &b->voice[voice];
	if (v->ended) { crb_set_error("voice %zu has already ended", voice); return CRB200_E_ARGUMENT; }
	if ((rc = voice_reserve(b, v, frames)) != 0) return rc;
	memcpy(LIVE(b, v) + v->frames * b->channels, input, frames * b->channels * sizeof(cc_s16l));
	v->frames += frames;
	v->pushed += frames;
	return CRB200_OK;
}

int ClownResamplerB200_VoiceBatchEnd(ClownResamplerB200_VoiceBatch *b, size_t voice)
{
	crb_voice *v;
	int rc;
	if (!b || voice >= b->voices) { crb_set_error("bad argument"); return CRB200_E_ARGUMENT; }
	v = &b->voice[voice];
	if (v->ended) return CRB200_OK;
	/* H:1223-1233: R frames of silence flush the tail */
	if ((rc = voice_reserve(b, v, b->radius)) != 0) return rc;
	memset(LIVE(b, v) + v->frames * b->channels, 0, b->radius * b->channels * sizeof(cc_s16l));
	v->frames += b->radius;
	v->ended = 1;
	return CRB200_OK;
}

int ClownResamplerB200_VoiceBatchAdjust(ClownResamplerB200_VoiceBatch *b, size_t voice,
	cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate)
{
	/* H:1183-1209 for one voice: the new rates take effect at the next output frame; the position is kept.  Refused, as
	   the reference refuses them, when the new kernel is wider than the one the voice was created with (H:1195) or does
	   not fit the wrapper's buffer (H:1202). */
	ClownResampler_LowLevel_State st;
	int plan;
	if (!b || voice >= b->voices) { crb_set_error("bad argument"); return CRB200_E_ARGUMENT; }
	if (b->in_flight) { crb_set_error("VoiceBatchAdjust between TickBegin and TickEnd"); return CRB200_E_ARGUMENT; }
	st = b->init;
	if (!ClownResampler_LowLevel_Adjust(&st, input_sample_rate, output_sample_rate, low_pass_filter_sample_rate)
	    || st.lowest_level.integer_stretched_kernel_radius > b->radius
	    || st.lowest_level.integer_stretched_kernel_radius * 2 >= HL_SAMPLES / b->channels) {
		crb_set_error("VoiceBatchAdjust: rates %lu -> %lu (low-pass %lu) need a kernel radius beyond the one the batch was created with (%zu frames); "
			"ClownResampler_HighLevel_Adjust refuses them too (H:1195, H:1202)", input_sample_rate, output_sample_rate, low_pass_filter_sample_rate, b->radius);
		return CRB200_E_CONFIG;
	}
	if (st.increment == 0 || st.increment > 0xFFFFFFFFul) { crb_set_error("VoiceBatchAdjust: bad ratio"); return CRB200_E_CONFIG; }
	if ((plan = plan_for(b, &st.lowest_level, st.increment)) < 0) return CRB200_E_CONFIG;
	b->voice[voice].increment = st.increment;
	b->voice[voice].radius = (unsigned)st.lowest_level.integer_stretched_kernel_radius;
	b->voice[voice].plan = plan;
	return CRB200_OK;
}

static int staging_reserve(ClownResamplerB200_VoiceBatch *b, size_t in_bytes, size_t out_bytes)
{
	if (in_bytes > b->in_cap) {
		size_t cap = b->in_cap ? b->in_cap : (1u << 20);
		while (cap < in_bytes) cap *= 2;
		crb_dev_pinned_free(b->pin_in); crb_dev_free(b->dev_in);
		b->pin_in = (cc_s16l *)crb_dev_pinned_alloc(cap); b->dev_in = crb_dev_alloc(cap + 64);
		b->in_cap = (b->pin_in && b->dev_in) ? cap : 0;
		if (!b->in_cap) return CRB200_E_MEMORY;
	}
	if (out_bytes > b->out_cap) {
		size_t cap = b->out_cap ? b->out_cap : (1u << 20);
		while (cap < out_bytes) cap *= 2;
		crb_dev_pinned_free(b->pin_out); crb_dev_free(b->dev_out);
		b->pin_out = (unsigned char *)crb_dev_pinned_alloc(cap); b->dev_out = crb_dev_alloc(cap);
		b->out_cap = (b->pin_out && b->dev_out) ? cap : 0;
		if (!b->out_cap) return CRB200_E_MEMORY;
	}
	return 0;
}

/* One tick, first half: how many frames every voice can emit, gather their input slices into pinned memory, one upload, one
   launch per kernel geometry in use, one download -- all asynchronous on the batch's stream. */
static int tick_begin(ClownResamplerB200_VoiceBatch *b, size_t max_frames, int output_format,
	void *output, size_t output_stride_bytes, size_t *produced)
{
	const size_t ch = b->channels, R = b->radius;
	const size_t fb_out = output_format == CRB200_OUT_S16_CLAMPED ? 2 * ch : 4 * ch;
	size_t i, in_bytes_total = 0, out_frames_total = 0, jobs_off = 0, download_bytes = 0, n_active = 0;
	size_t in_off = 0, out_off = 0, n_jobs = 0;
	size_t plan_first_job[VOICE_PLANS + 1];
	uint64_t plan_tiles[VOICE_PLANS];
	crb_device_job *pinned_jobs;
	int rc, direct, k;
	double t0 = 0, t1 = 0, t2 = 0;

	if (b->trace) t0 = crb_now();
	/* 1. how many frames can every voice emit, and which slice of its padded stream do they read */
	for (i = 0; i < b->voices; ++i) {
		crb_voice *v = &b->voice[i];
		/* frames usable as input: everything pushed once ended; otherwise the last R frames are only look-ahead
		   (H:1143-1154: the second dead zone) */
		const size_t total = v->ended ? v->pushed : (v->pushed > R ? v->pushed - R : 0);
		ClownResampler_LowLevel_State at;
		size_t n;
		at.position_integer = v->pos_int;
		at.position_fractional = v->pos_frac;
		at.increment = v->increment;
		n = ClownResamplerB200_CountOutputFrames(&at, total);       /* frames H:1058-1092 would still emit from here */
		if (n > max_frames) n = max_frames;
		produced[i] = n;
		if (n) {
			const u128 p0 = ((u128)v->pos_int << 16) + v->pos_frac, p1 = p0 + (u128)(n - 1) * v->increment;
			/* a voice of radius r <= R hands the low-level loop the buffer r frames before its position (H:1168): its window
			   base sits R - r frames further into the R-padded stream */
			const size_t shift = R - v->radius;
			crb_voice_slice *sl = &b->slice[i];
			sl->first = (size_t)(p0 >> 16) + shift;
			sl->last = (size_t)(p1 >> 16) + shift + 2 * v->radius + 1;            /* exclusive */
			if (sl->last > v->base + v->frames) sl->last = v->base + v->frames;
			/* position of the voice's next frame relative to the slice (the plan's radius delta is added per plan below) */
			sl->q0 = (uint64_t)(p0 + ((u128)shift << 16) - ((u128)sl->first << 16));
			in_bytes_total = ((in_bytes_total + 15) & ~(size_t)15) + (sl->last - sl->first) * ch * 2;   /* slices start 16-byte aligned */
			out_frames_total += n;
			++n_active;
		}
	}
	if (out_frames_total == 0) return CRB200_OK;
	/* a pinned caller buffer takes the download directly (voice v at its stride, no scatter copy); the job tables
	   travel behind the input slices in the same upload */
	direct = output_stride_bytes >= max_frames * fb_out && output_stride_bytes % 16 == 0
		&& crb_dev_is_pinned(output, b->voices * output_stride_bytes);
	jobs_off = (in_bytes_total + 63) & ~(size_t)63;
	if ((rc = staging_reserve(b, jobs_off + n_active * sizeof(crb_device_job) + 64,
	                          direct ? b->voices * output_stride_bytes : out_frames_total * fb_out)) != 0) return rc;
	pinned_jobs = (crb_device_job *)((unsigned char *)b->pin_in + jobs_off);

	if (b->trace) t1 = crb_now();
	/* 2. gather the slices into pinned memory, plan by plan; consecutive voices of an unstretched plan that walk through the same
	      phases (same position fraction, frame count and step: voices started together) become one lockstep job of 4 or 2 */
	for (k = 0; k < b->n_plans; ++k) {
		const ClownResamplerB200_Plan *plan = b->plans[k].plan;
		const uint32_t tile_out = plan->geo.tile_out;
		const int can_lock = plan->kernel_kind == 0 && plan->geo.unstretched5;
		uint64_t tiles = 0;
		size_t run_start = (size_t)-1, run_len = 0, prev_i = 0;
		plan_first_job[k] = n_jobs;
		for (i = 0; i <= b->voices; ++i) {
			const int active = i < b->voices && produced[i] && b->voice[i].plan == k;
			int same = 0;
			if (i < b->voices && !active) continue;
			if (active) {
				crb_voice *v = &b->voice[i];
				crb_voice_slice *sl = &b->slice[i];
				const size_t bytes = (sl->last - sl->first) * ch * 2;
				in_off = (in_off + 15) & ~(size_t)15;
				memcpy((unsigned char *)b->pin_in + in_off, LIVE(b, v) + (sl->first - v->base) * ch, bytes);
				sl->dev_in = (const int16_t *)((unsigned char *)b->dev_in + in_off);
				sl->dev_out = (unsigned char *)b->dev_out + (direct ? i * output_stride_bytes : out_off);
				if (direct) { if (i * output_stride_bytes + produced[i] * fb_out > download_bytes) download_bytes = i * output_stride_bytes + produced[i] * fb_out; }
				else download_bytes = out_off + produced[i] * fb_out;
				in_off += bytes;
				out_off += produced[i] * fb_out;
				same = can_lock && run_len > 0 && run_len < CRB_MAX_LOCKSTEP && b->slice[prev_i].q0 == sl->q0 && produced[prev_i] == produced[i]
					&& b->voice[prev_i].increment == v->increment;
			}
			if (same) {
				++run_len;
				prev_i = i;
				continue;
			}
			/* flush the run [run_start ...] of run_len lockstep voices as jobs of 4, 2 or 1 streams */
			if (run_len) {
				size_t members[CRB_MAX_LOCKSTEP], m = 0, j, at = 0;
				for (j = run_start; m < run_len; ++j)
					if (j < b->voices && produced[j] && b->voice[j].plan == k) members[m++] = j;
				while (at < run_len) {
					const size_t left = run_len - at;
					const unsigned log_streams = left >= 4 && plan->geo.lock_slot_bytes[2] ? 2 : left >= 2 && plan->geo.lock_slot_bytes[1] ? 1 : 0;
					const size_t width = (size_t)1 << log_streams;
					crb_device_job *job = &pinned_jobs[n_jobs++];
					const crb_voice_slice *s0 = &b->slice[members[at]];
					memset(job, 0, sizeof *job);
					job->in = s0->dev_in;
					job->out = s0->dev_out;
					job->q0 = s0->q0 + plan->geo.delta;
					job->n_out = produced[members[at]];
					job->in_frames = s0->last - s0->first;
					job->increment = b->voice[members[at]].increment;
					job->n_more = (uint32_t)width - 1;
					for (j = 1; j < width; ++j) {
						const crb_voice_slice *sj = &b->slice[members[at + j]];
						job->in_more[j - 1] = sj->dev_in;
						job->out_more[j - 1] = sj->dev_out;
						job->in_frames_more[j - 1] = sj->last - sj->first;
					}
					job->tile_base = tiles;
					tiles += (job->n_out + (tile_out >> log_streams) - 1) / (tile_out >> log_streams);
					at += width;
				}
			}
			run_start = i;
			run_len = active ? 1 : 0;
			prev_i = i;
		}
		plan_tiles[k] = tiles;
	}
	plan_first_job[b->n_plans] = n_jobs;
	if (b->trace) t2 = crb_now();
	/* 3. one upload, one launch per geometry, one download */
	if ((rc = crb_dev_h2d(b->dev_in, b->pin_in, jobs_off + n_jobs * sizeof(crb_device_job), b->stream)) != 0) return rc;
	for (k = 0; k < b->n_plans; ++k) {
		const size_t nj = plan_first_job[k + 1] - plan_first_job[k];
		if (!nj) continue;
		if ((rc = crb_dev_launch_resident(b->plans[k].plan, (const crb_device_job *)((unsigned char *)b->dev_in + jobs_off) + plan_first_job[k], nj,
		                                  plan_tiles[k], output_format, b->stream)) != 0) return rc;
	}
	if ((rc = crb_dev_d2h(direct ? output : (void *)b->pin_out, b->dev_out, download_bytes, b->stream)) != 0) return rc;
	b->in_flight = 1; b->direct = direct; b->fmt = output_format;
	b->out = output; b->out_stride = output_stride_bytes; b->produced = produced;
	if (b->trace) { const double t3 = crb_now(); b->t_plan += t1 - t0; b->t_gather += t3 - t1; (void)t2; }
	return CRB200_OK;
}

/* second half: wait for the download, scatter the frames, advance the voices, drop input that no later frame can read */
static int tick_end(ClownResamplerB200_VoiceBatch *b)
{
	const size_t ch = b->channels;
	const size_t fb_out = b->fmt == CRB200_OUT_S16_CLAMPED ? 2 * ch : 4 * ch;
	size_t i, out_off = 0;
	int rc, k;
	double t0 = 0, t1 = 0;
	if (!b->in_flight) return CRB200_OK;
	if (b->trace) t0 = crb_now();
	rc = crb_dev_sync(b->stream);
	b->in_flight = 0;
	if (rc != 0) return rc;
	if (b->trace) t1 = crb_now();
	/* the staging order was by plan, then by voice */
	for (k = 0; k < b->n_plans; ++k)
		for (i = 0; i < b->voices; ++i) {
			crb_voice *v = &b->voice[i];
			const size_t n = b->produced[i];
			size_t keep_from;
			if (!n || v->plan != k) continue;
			if (!b->direct) memcpy((unsigned char *)b->out + i * b->out_stride, b->pin_out + out_off, n * fb_out);
			out_off += n * fb_out;
			{
				const u128 next = ((u128)v->pos_int << 16) + v->pos_frac + (u128)n * v->increment;
				v->pos_int = (size_t)(next >> 16);
				v->pos_frac = (cc_u32f)(next & 0xFFFF);
			}
			keep_from = v->pos_int;      /* first frame any later window can touch (a window never starts before its position) */
			if (keep_from > v->base + v->frames) keep_from = v->base + v->frames;
			if (keep_from > v->base) {
				const size_t drop = keep_from - v->base;
				v->off += drop;
				v->frames -= drop;
				v->base += drop;
			}
		}
	if (b->trace) {
		const double t2 = crb_now();
		b->t_device += t1 - t0; b->t_scatter += t2 - t1; ++b->ticks;
	}
	return CRB200_OK;
}

int ClownResamplerB200_VoiceBatchTickBegin(ClownResamplerB200_VoiceBatch *b, size_t max_frames, int output_format,
	void *output, size_t output_stride_bytes, size_t *produced)
{
	int rc, prev;
	if (!b || !output || !produced || (output_format != CRB200_OUT_S32 && output_format != CRB200_OUT_S16_CLAMPED)) { crb_set_error("bad argument"); return CRB200_E_ARGUMENT; }
	if (b->in_flight) { crb_set_error("VoiceBatchTickBegin: the previous tick has not been ended"); return CRB200_E_ARGUMENT; }
	prev = crb_dev_push(b->device);
	rc = tick_begin(b, max_frames, output_format, output, output_stride_bytes, produced);
	crb_dev_pop(prev);
	return rc;
}

int ClownResamplerB200_VoiceBatchTickEnd(ClownResamplerB200_VoiceBatch *b)
{
	int rc, prev;
	if (!b) { crb_set_error("bad argument"); return CRB200_E_ARGUMENT; }
	prev = crb_dev_push(b->device);
	rc = tick_end(b);
	crb_dev_pop(prev);
	return rc;
}

int ClownResamplerB200_VoiceBatchTick(ClownResamplerB200_VoiceBatch *b, size_t max_frames, int output_format,
	void *output, size_t output_stride_bytes, size_t *produced)
{
	const int rc = ClownResamplerB200_VoiceBatchTickBegin(b, max_frames, output_format, output, output_stride_bytes, produced);
	return rc != CRB200_OK ? rc : ClownResamplerB200_VoiceBatchTickEnd(b);
}
