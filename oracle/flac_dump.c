/*
 * flac_dump.c -- decodes the reference's FLAC fixture to raw interleaved s16le with the
 * reference's own vendored decoder (tests/dr_flac.h, included by path, exactly like
 * tests/test-low-level.c:93,148 does).  Used once by oracle/make_golden.py to produce
 * tests/golden/test_flac_s16le.bin.gz.  TEST INFRASTRUCTURE ONLY.
 */
#define DR_FLAC_IMPLEMENTATION
#include <dr_flac.h>
#include <stdio.h>
#include <stdlib.h>

int main(int argc, char **argv)
{
    drflac *f;
    drflac_int16 *pcm;
    drflac_uint64 n;
    FILE *o;
    if (argc < 3) { fprintf(stderr, "usage: %s in.flac out.s16\n", argv[0]); return 1; }
    f = drflac_open_file(argv[1], NULL);
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    pcm = (drflac_int16 *)malloc((size_t)f->totalPCMFrameCount * f->channels * sizeof(drflac_int16));
    n = drflac_read_pcm_frames_s16(f, f->totalPCMFrameCount, pcm);
    fprintf(stderr, "channels=%u rate=%u bits=%u frames=%llu\n", f->channels, f->sampleRate, f->bitsPerSample, (unsigned long long)n);
    o = fopen(argv[2], "wb");
    fwrite(pcm, sizeof(drflac_int16) * f->channels, (size_t)n, o);
    fclose(o);
    drflac_close(f);
    free(pcm);
    return 0;
}
