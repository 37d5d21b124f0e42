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
 */
#include "../../include/clownresampler_b200.h"
#include "crb_internal.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef unsigned __int128 u128;

/* CRB200_TRACE=1: per-phase wall time of VoiceBatchTick, printed when the batch is destroyed */
static double crb_now(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

typedef struct crb_voice {
	cc_s16l *data;          /* padded stream from frame `base` on: R zeros, pushed input, (R zeros once ended) */
	size_t base;            /* padded-stream index of data[0] */
	size_t frames;          /* frames stored in data */
	size_t capacity;        /* frames allocated */
	size_t pushed;          /* input frames pushed so far */
	size_t pos_int;         /* position of the next output frame in the padded stream (H:645-646) */
	cc_u32f pos_frac;
	cc_u32f increment;      /* this voice's 16.16 step (H:647); changed by VoiceBatchAdjust */
	int ended;
} crb_voice;

struct ClownResamplerB200_VoiceBatch {
	ClownResamplerB200_Plan *plan;          /* built for the largest increment of any voice (tile sizing) */
	int device;                             /* the device the plan, the stream and the staging buffers live on */
	ClownResampler_Precomputed *table;      /* copy of the caller's table, for re-planning after an Adjust */
	ClownResampler_LowLevel_State init;     /* configuration shared by all voices; increment = the plan's */
	size_t voices, channels, radius;
	crb_voice *voice;
	/* per-tick staging */
	void *stream;
	cc_s16l *pin_in; void *dev_in; size_t in_cap;
	unsigned char *pin_out; void *dev_out; size_t out_cap;
	size_t *slice_first;
	int trace; double t_plan, t_gather, t_device, t_scatter; size_t ticks;
};

static int voice_reserve(ClownResamplerB200_VoiceBatch *b, crb_voice *v, size_t extra)
{
	if (v->frames + extra > v->capacity) {
		size_t cap = v->capacity ? v->capacity : 1024;
		cc_s16l *p;
		while (cap < v->frames + extra) cap *= 2;
		p = (cc_s16l *)realloc(v->data, cap * b->channels * sizeof(cc_s16l));
		if (!p) { crb_set_error("out of host memory"); return CRB200_E_MEMORY; }
		v->data = p;
		v->capacity = cap;
	}
	return 0;
}

ClownResamplerB200_VoiceBatch *ClownResamplerB200_VoiceBatchCreate(const ClownResampler_Precomputed *precomputed,
	size_t voices, cc_u8f channels, cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate)
{
	ClownResamplerB200_VoiceBatch *b;
	size_t i;
	int prev;
	if (!precomputed || voices == 0) { crb_set_error("bad argument"); return NULL; }
	b = (ClownResamplerB200_VoiceBatch *)calloc(1, sizeof *b);
	if (!b) { crb_set_error("out of host memory"); return NULL; }
	if (channels == 0 || channels > CLOWNRESAMPLER_MAXIMUM_CHANNELS                                                /* H:1103 */
	    || !ClownResampler_LowLevel_Init(&b->init, channels, input_sample_rate, output_sample_rate, low_pass_filter_sample_rate)) {
		crb_set_error("configuration rejected (channels %u, rates %lu -> %lu, low-pass %lu)", channels, input_sample_rate, output_sample_rate, low_pass_filter_sample_rate);
		free(b);
		return NULL;
	}
	b->plan = ClownResamplerB200_PlanCreate(precomputed, &b->init);
	b->device = b->plan ? b->plan->device : -1;
	prev = b->plan ? crb_dev_push(b->device) : -1;
	b->table = (ClownResampler_Precomputed *)malloc(sizeof *b->table);
	if (b->table) *b->table = *precomputed;
	b->voice = (crb_voice *)calloc(voices, sizeof *b->voice);
	b->slice_first = (size_t *)malloc(voices * sizeof *b->slice_first);
	b->stream = b->plan ? crb_dev_stream_create() : NULL;
	crb_dev_pop(prev);
	if (!b->plan || !b->table || !b->voice || !b->slice_first || !b->stream) {
		if (b->plan && (!b->table || !b->voice || !b->slice_first)) crb_set_error("out of host memory");
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
		memset(v->data, 0, b->radius * channels * sizeof(cc_s16l));
		v->frames = b->radius;
		v->increment = b->init.increment;
	}
	return b;
}

void ClownResamplerB200_VoiceBatchDestroy(ClownResamplerB200_VoiceBatch *b)
{
	size_t i;
	if (!b) return;
	if (b->trace && b->ticks)
		fprintf(stderr, "clownresampler_b200 VoiceBatch: %zu ticks; per tick: plan %.1f us, gather %.1f us, upload+kernel+download %.1f us, scatter %.1f us\n",
			b->ticks, 1e6 * b->t_plan / b->ticks, 1e6 * b->t_gather / b->ticks, 1e6 * b->t_device / b->ticks, 1e6 * b->t_scatter / b->ticks);
	if (b->voice) for (i = 0; i < b->voices; ++i) free(b->voice[i].data);
	free(b->voice); free(b->slice_first); free(b->table);
	{
		const int prev = b->plan ? crb_dev_push(b->device) : -1;
		crb_dev_pinned_free(b->pin_in); crb_dev_free(b->dev_in);
		crb_dev_pinned_free(b->pin_out); crb_dev_free(b->dev_out);
		crb_dev_stream_destroy(b->stream);
		crb_dev_pop(prev);
	}
	if (b->plan) ClownResamplerB200_PlanDestroy(b->plan);
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
	memcpy(v->data + v->frames * b->channels, input, frames * b->channels * sizeof(cc_s16l));
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
	memset(v->data + v->frames * b->channels, 0, b->radius * b->channels * sizeof(cc_s16l));
	v->frames += b->radius;
	v->ended = 1;
	return CRB200_OK;
}

int ClownResamplerB200_VoiceBatchAdjust(ClownResamplerB200_VoiceBatch *b, size_t voice,
	cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate)
{
	/* H:1183-1209 for one voice: the new rates take effect at the next output frame; the position is kept.  The
	   voices of a batch share one kernel geometry (H:632-638), so the new rates must give the same one: any
	   up-sampling ratio in an unstretched batch, or the same low-pass scale otherwise. */
	ClownResampler_LowLevel_State st;
	if (!b || voice >= b->voices) { crb_set_error("bad argument"); return CRB200_E_ARGUMENT; }
	st = b->init;
	if (!ClownResampler_LowLevel_Adjust(&st, input_sample_rate, output_sample_rate, low_pass_filter_sample_rate)
	    || memcmp(&st.lowest_level, &b->init.lowest_level, sizeof st.lowest_level) != 0) {
		crb_set_error("VoiceBatchAdjust: rates %lu -> %lu (low-pass %lu) need a different kernel geometry than the batch was created with",
			input_sample_rate, output_sample_rate, low_pass_filter_sample_rate);
		return CRB200_E_CONFIG;
	}
	if (st.increment == 0 || st.increment > 0xFFFFFFFFul) { crb_set_error("VoiceBatchAdjust: bad ratio"); return CRB200_E_CONFIG; }
	b->voice[voice].increment = st.increment;
	return CRB200_OK;
}

/* the plan's tiles are sized for its increment: re-plan when some voice now steps faster */
static int replan_if_needed(ClownResamplerB200_VoiceBatch *b)
{
	cc_u32f largest = 0;
	size_t i;
	for (i = 0; i < b->voices; ++i) if (b->voice[i].increment > largest) largest = b->voice[i].increment;
	if (largest > b->init.increment) {
		ClownResampler_LowLevel_State st = b->init;
		ClownResamplerB200_Plan *plan;
		st.increment = largest;
		plan = ClownResamplerB200_PlanCreateOnDevice(b->table, &st, b->device);
		if (!plan) return CRB200_E_CONFIG;
		ClownResamplerB200_PlanDestroy(b->plan);
		b->plan = plan;
		b->init.increment = largest;
	}
	return 0;
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

static int voice_batch_tick(ClownResamplerB200_VoiceBatch *b, size_t max_frames, int output_format,
	void *output, size_t output_stride_bytes, size_t *produced);

int ClownResamplerB200_VoiceBatchTick(ClownResamplerB200_VoiceBatch *b, size_t max_frames, int output_format,
	void *output, size_t output_stride_bytes, size_t *produced)
{
	int rc, prev;
	if (!b) { crb_set_error("bad argument"); return CRB200_E_ARGUMENT; }
	prev = crb_dev_push(b->device);
	rc = voice_batch_tick(b, max_frames, output_format, output, output_stride_bytes, produced);
	crb_dev_pop(prev);
	return rc;
}

static int voice_batch_tick(ClownResamplerB200_VoiceBatch *b, size_t max_frames, int output_format,
	void *output, size_t output_stride_bytes, size_t *produced)
{
	const size_t ch = b ? b->channels : 0, R = b ? b->radius : 0;
	const size_t fb_out = output_format == CRB200_OUT_S16_CLAMPED ? 2 * ch : 4 * ch;
	size_t i, in_bytes_total = 0, n_jobs = 0, out_frames_total = 0, jobs_off = 0, download_bytes = 0;
	crb_device_job *pinned_jobs = NULL;
	uint64_t tiles = 0;
	int rc, direct = 0;
	double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
	if (!b || !output || !produced || (output_format != CRB200_OUT_S32 && output_format != CRB200_OUT_S16_CLAMPED)) { crb_set_error("bad argument"); return CRB200_E_ARGUMENT; }

	if ((rc = replan_if_needed(b)) != 0) return rc;
	if (b->trace) t0 = crb_now();
	/* 1. how many frames can every voice emit, and which slice of its padded stream do they read */
	for (i = 0; i < b->voices; ++i) {
		crb_voice *v = &b->voice[i];
		/* frames usable as input: everything pushed once ended; otherwise the last R frames are only look-ahead
		   (H:1143-1154: the second dead zone) */
		const size_t total = v->ended ? v->pushed : (v->pushed > R ? v->pushed - R : 0);
		ClownResampler_LowLevel_State at = b->init;
		size_t n;
		at.position_integer = v->pos_int;
		at.position_fractional = v->pos_frac;
		at.increment = v->increment;
		n = ClownResamplerB200_CountOutputFrames(&at, total);       /* frames H:1058-1092 would still emit from here */
		if (n > max_frames) n = max_frames;
		produced[i] = n;
		if (n) {
			const u128 p0 = ((u128)v->pos_int << 16) + v->pos_frac, p1 = p0 + (u128)(n - 1) * v->increment;
			const size_t first = (size_t)(p0 >> 16);                 /* padded-stream frame of the first window base */
			size_t last = (size_t)(p1 >> 16) + 2 * R + 1;            /* exclusive */
			if (last > v->base + v->frames) last = v->base + v->frames;
			b->slice_first[i] = first;
			in_bytes_total = ((in_bytes_total + 15) & ~(size_t)15) + (last - first) * ch * 2;   /* slices start 16-byte aligned */
			out_frames_total += n;
		}
	}
	if (out_frames_total == 0) return CRB200_OK;
	/* a pinned caller buffer takes the download directly (voice v at its stride, no scatter copy); the job table
	   travels behind the input slices in the same upload */
	direct = output_stride_bytes >= max_frames * fb_out && output_stride_bytes % 16 == 0
		&& crb_dev_is_pinned(output, b->voices * output_stride_bytes);
	jobs_off = (in_bytes_total + 63) & ~(size_t)63;
	if ((rc = staging_reserve(b, jobs_off + b->voices * sizeof(crb_device_job) + 64,
	                          direct ? b->voices * output_stride_bytes : out_frames_total * fb_out)) != 0) return rc;
	pinned_jobs = (crb_device_job *)((unsigned char *)b->pin_in + jobs_off);

	if (b->trace) t1 = crb_now();
	/* 2. gather the slices into pinned memory, one job per active voice */
	{
		size_t in_off = 0, out_off = 0;       /* bytes */
		for (i = 0; i < b->voices; ++i) {
			crb_voice *v = &b->voice[i];
			const size_t n = produced[i];
			size_t first, last, bytes;
			u128 p0;
			crb_device_job *j;
			if (!n) continue;
			first = b->slice_first[i];
			p0 = ((u128)v->pos_int << 16) + v->pos_frac;
			last = (size_t)((p0 + (u128)(n - 1) * v->increment) >> 16) + 2 * R + 1;
			if (last > v->base + v->frames) last = v->base + v->frames;
			in_off = (in_off + 15) & ~(size_t)15;
			bytes = (last - first) * ch * 2;
			memcpy((unsigned char *)b->pin_in + in_off, v->data + (first - v->base) * ch, bytes);
			j = &pinned_jobs[n_jobs++];
			memset(j, 0, sizeof *j);
			j->in = (const int16_t *)((unsigned char *)b->dev_in + in_off);
			j->out = (unsigned char *)b->dev_out + (direct ? i * output_stride_bytes : out_off);
			download_bytes = direct ? i * output_stride_bytes + n * fb_out : out_off + n * fb_out;
			j->q0 = (uint64_t)(p0 - ((u128)first << 16)) + b->plan->geo.delta;
			j->first_out = 0;
			j->n_out = n;
			j->in_frames = last - first;
			j->increment = v->increment;
			j->tile_base = tiles;
			tiles += (n + b->plan->geo.tile_out - 1) / b->plan->geo.tile_out;
			in_off += bytes;
			out_off += n * fb_out;
		}
		if (b->trace) t2 = crb_now();
		/* 3. one upload, one launch, one download */
		if ((rc = crb_dev_h2d(b->dev_in, b->pin_in, jobs_off + n_jobs * sizeof(crb_device_job), b->stream)) != 0) return rc;
		if ((rc = crb_dev_launch_resident(b->plan, (const crb_device_job *)((unsigned char *)b->dev_in + jobs_off), n_jobs, tiles, output_format, b->stream)) != 0) return rc;
		if ((rc = crb_dev_d2h(direct ? output : (void *)b->pin_out, b->dev_out, download_bytes, b->stream)) != 0) return rc;
		if ((rc = crb_dev_sync(b->stream)) != 0) return rc;
	}

	if (b->trace) t3 = crb_now();
	/* 4. scatter the frames, advance the voices, drop input that no later frame can read */
	{
		size_t out_off = 0;
		for (i = 0; i < b->voices; ++i) {
			crb_voice *v = &b->voice[i];
			const size_t n = produced[i];
			size_t keep_from;
			if (!n) continue;
			if (!direct) memcpy((unsigned char *)output + i * output_stride_bytes, b->pin_out + out_off, n * fb_out);
			out_off += n * fb_out;
			{
				const u128 next = ((u128)v->pos_int << 16) + v->pos_frac + (u128)n * v->increment;
				v->pos_int = (size_t)(next >> 16);
				v->pos_frac = (cc_u32f)(next & 0xFFFF);
			}
			keep_from = v->pos_int;                                  /* first frame the next output frame can touch */
			if (keep_from > v->base + v->frames) keep_from = v->base + v->frames;
			if (keep_from > v->base) {
				const size_t drop = keep_from - v->base;
				memmove(v->data, v->data + drop * ch, (v->frames - drop) * ch * sizeof(cc_s16l));
				v->frames -= drop;
				v->base += drop;
			}
		}
	}
	if (b->trace) {
		const double t4 = crb_now();
		b->t_plan += t1 - t0; b->t_gather += t2 - t1; b->t_device += t3 - t2; b->t_scatter += t4 - t3; ++b->ticks;
	}
	return CRB200_OK;
}
