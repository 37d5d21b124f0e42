/*
 * bench_highlevel.c -- config 4 of BASELINE.json: many mono game-audio voices 22.05 kHz -> 48 kHz through the
 * HighLevel input/output-callback API, one 1024-frame tick at a time.
 *
 * The same source builds twice:
 *   cc -DUSE_REFERENCE -I/root/reference ...         the unmodified reference (header-only, static)
 *   cc -Iinclude ... -lclownresampler_b200           the drop-in library (+ -DWITH_BATCH for the VoiceBatch path)
 * usage: bench_highlevel voices seconds [batch]
 */
#ifdef USE_REFERENCE
#define CLOWNRESAMPLER_IMPLEMENTATION
#define CLOWNRESAMPLER_STATIC
#include <clownresampler.h>
#else
#include "clownresampler_b200.h"
#endif

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define TICK 1024

typedef struct voice {
	ClownResampler_HighLevel_State state;
	const cc_s16l *data;
	size_t frames, pos;
	cc_s16l *out;
	size_t out_pos, tick_left;
	int finished, ended_input, ended_at_begin, drained;
} voice;

static size_t in_cb(void *user, cc_s16l *buffer, size_t total_frames)
{
	voice *v = (voice *)user;
	size_t n = v->frames - v->pos;
	if (n > total_frames) n = total_frames;
	memcpy(buffer, v->data + v->pos, n * sizeof(cc_s16l));
	v->pos += n;
	return n;
}

static cc_bool out_cb(void *user, const cc_s32f *frame, cc_u8f n)
{
	voice *v = (voice *)user;
	const cc_s32f s = frame[0];
	(void)n;
	v->out[v->out_pos++] = (cc_s16l)(s < -0x7FFF ? -0x7FFF : (s > 0x7FFF ? 0x7FFF : s));   /* examples/high-level.c clamp */
	return --v->tick_left != 0;
}

static double now(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

int main(int argc, char **argv)
{
	const size_t n_voices = argc > 1 ? (size_t)atol(argv[1]) : 16;
	const size_t seconds = argc > 2 ? (size_t)atol(argv[2]) : 10;
	const int batch = argc > 3;
	const size_t T = 22050 * seconds, out_cap = T * 48000 / 22050 + 4096;
	static ClownResampler_Precomputed pre;
	cc_s16l *input = (cc_s16l *)malloc(T * sizeof(cc_s16l));
	voice *voices = (voice *)calloc(n_voices, sizeof(voice));
	size_t i, v, ticks = 0, total_out = 0;
	unsigned long checksum = 0;
	unsigned x = 12345;
	double t0, t1;

	for (i = 0; i < T; ++i) { x = x * 1664525u + 1013904223u; input[i] = (cc_s16l)(x >> 16); }
	ClownResampler_Precompute(&pre);
	for (v = 0; v < n_voices; ++v) {
		ClownResampler_HighLevel_Init(&voices[v].state, 1, 22050, 48000, 48000);
		voices[v].data = input; voices[v].frames = T;
		voices[v].out = (cc_s16l *)malloc(out_cap * sizeof(cc_s16l));
	}

#ifndef USE_REFERENCE
	if (ClownResamplerB200_Init(0) != 0) { fprintf(stderr, "%s\n", ClownResamplerB200_GetLastError()); return 1; }   /* context creation is not part of a tick */
#endif
	t0 = now();
	if (!batch) {
		/* every voice is advanced by its own HighLevel_Resample call per tick, as a game mixer would */
		int active = 1;
		while (active) {
			active = 0;
			for (v = 0; v < n_voices; ++v) {
				voice *vc = &voices[v];
				if (vc->finished) continue;
				vc->tick_left = TICK;
				if (ClownResampler_HighLevel_Resample(&vc->state, &pre, in_cb, out_cb, vc)) {
					if (vc->tick_left != 0 && ClownResampler_HighLevel_ResampleEnd(&vc->state, &pre, out_cb, vc))
						vc->finished = 1;
				}
				active |= !vc->finished;
			}
			++ticks;
		}
	}
#ifdef WITH_BATCH
	else {
		ClownResamplerB200_VoiceBatch *b = ClownResamplerB200_VoiceBatchCreate(&pre, n_voices, 1, 22050, 48000, 48000);
		size_t *produced = (size_t *)malloc(n_voices * sizeof(size_t));
		cc_s16l *tick_out = (cc_s16l *)ClownResamplerB200_PinnedAlloc(n_voices * TICK * sizeof(cc_s16l));   /* pinned: the download lands here directly */
		const size_t per_tick_in = (size_t)((double)TICK * 22050.0 / 48000.0) + 2;
		size_t done = 0;
		if (!b) { fprintf(stderr, "%s\n", ClownResamplerB200_GetLastError()); return 1; }
		/* double-buffered ticks: while the GPU works on tick n (TickBegin .. TickEnd) the host consumes the output of tick n - 1
		   and pushes the input of tick n + 1 -- what a mixer thread with a decoder beside it does */
		cc_s16l *tick_out2 = (cc_s16l *)ClownResamplerB200_PinnedAlloc(n_voices * TICK * sizeof(cc_s16l));
		size_t *produced2 = (size_t *)malloc(n_voices * sizeof(size_t));
		cc_s16l *bufs[2]; size_t *prods[2];
		int cur = 0, have_prev = 0;
		bufs[0] = tick_out; bufs[1] = tick_out2; prods[0] = produced; prods[1] = produced2;
#define PUSH_NEXT() do { \
			for (v = 0; v < n_voices; ++v) { \
				voice *vc = &voices[v]; \
				size_t n = vc->frames - vc->pos; \
				if (vc->ended_input) continue; \
				if (n > per_tick_in) n = per_tick_in; \
				ClownResamplerB200_VoiceBatchPush(b, v, vc->data + vc->pos, n); \
				vc->pos += n; \
				if (vc->pos == vc->frames) { ClownResamplerB200_VoiceBatchEnd(b, v); vc->ended_input = 1; } \
			} } while (0)
		PUSH_NEXT();
		while (done < n_voices) {
			if (ClownResamplerB200_VoiceBatchTickBegin(b, TICK, CRB200_OUT_S16_CLAMPED, bufs[cur], TICK * sizeof(cc_s16l), prods[cur]) != 0) {
				fprintf(stderr, "%s\n", ClownResamplerB200_GetLastError()); return 1;
			}
			if (have_prev) {
				/* consume the previous tick's output */
				for (v = 0; v < n_voices; ++v) {
					voice *vc = &voices[v];
					if (vc->finished) continue;
					memcpy(vc->out + vc->out_pos, bufs[cur ^ 1] + v * TICK, prods[cur ^ 1][v] * sizeof(cc_s16l));
					vc->out_pos += prods[cur ^ 1][v];
					if (vc->ended_input && vc->drained && prods[cur ^ 1][v] < TICK) { vc->finished = 1; ++done; }
				}
			}
			PUSH_NEXT();
			if (ClownResamplerB200_VoiceBatchTickEnd(b) != 0) { fprintf(stderr, "%s\n", ClownResamplerB200_GetLastError()); return 1; }
			/* a voice is drained when a tick that began after its input had ended produced less than a full tick */
			for (v = 0; v < n_voices; ++v) if (voices[v].ended_at_begin && prods[cur][v] < TICK) voices[v].drained = 1;
			for (v = 0; v < n_voices; ++v) voices[v].ended_at_begin = voices[v].ended_input;
			have_prev = 1;
			cur ^= 1;
			++ticks;
			if (done < n_voices) {
				int all = 1;
				for (v = 0; v < n_voices; ++v) if (!voices[v].drained) { all = 0; break; }
				if (all) {
					/* the last tick's output */
					for (v = 0; v < n_voices; ++v) {
						voice *vc = &voices[v];
						if (vc->finished) continue;
						memcpy(vc->out + vc->out_pos, bufs[cur ^ 1] + v * TICK, prods[cur ^ 1][v] * sizeof(cc_s16l));
						vc->out_pos += prods[cur ^ 1][v];
						vc->finished = 1; ++done;
					}
				}
			}
		}
#undef PUSH_NEXT
		ClownResamplerB200_VoiceBatchDestroy(b);
	}
#endif
	t1 = now();
	for (v = 0; v < n_voices; ++v) {
		total_out += voices[v].out_pos;
		for (i = 0; i < voices[v].out_pos; ++i) checksum = checksum * 31 + (unsigned short)voices[v].out[i];
	}
	printf("{\"impl\": \"%s\", \"voices\": %zu, \"seconds\": %zu, \"ticks\": %zu, \"output_frames\": %zu, \"wall_s\": %.4f, \"msamples_per_s\": %.2f, \"voice_ticks_per_s\": %.0f, \"checksum\": %lu}\n",
#ifdef USE_REFERENCE
	       "reference (CPU, 1 thread)",
#else
	       batch ? "b200 VoiceBatch" : "b200 drop-in HighLevel",
#endif
	       n_voices, seconds, ticks, total_out, t1 - t0, (double)total_out / (t1 - t0) / 1e6, (double)n_voices * (double)ticks / (t1 - t0), checksum);
	return 0;
}
