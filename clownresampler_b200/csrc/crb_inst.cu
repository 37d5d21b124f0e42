/*
 * crb_inst.cu -- instantiates crb_tiled_kernel<C, FMT, K> for one kernel kind K and one half of the channel
 * counts per translation unit (compiled once per (CRB_INST_K, CRB_INST_PART) by the Makefile, in parallel):
 *   PART 0: C = 0 (count at run time: 9..16 channels, and every count in the diagnostic format), 1, 2, 3, 4
 *   PART 1: C = 5, 6, 7, 8
 *   PART 2: C = 9 .. 12,  PART 3: C = 13 .. 16   (general and unstretched kinds only)
 * The slightly stretched kinds (K = 6, 8, 10, 12) exist for 1..8 channels only (plus C = 0 for the diagnostic format).
 */
#include "crb_kernels.cuh"

#ifndef CRB_INST_K
#error "compile with -DCRB_INST_K=<0|1|6|8|10|12> -DCRB_INST_PART=<0|1>"
#endif

#define CRB_CAT4(a, b, c, d) a##b##c##d
#define CRB_PICK_NAME(K, PART) CRB_CAT4(crb_pick_k, K, _p, PART)

template <int C, int FMT>
static crb_kernel_fn inst(unsigned *block)
{
	*block = CRB_NT_K(C, CRB_INST_K == 1) + 32;
	return (crb_kernel_fn)crb_tiled_kernel<C, FMT, CRB_INST_K>;
}

extern "C" crb_kernel_fn CRB_PICK_NAME(CRB_INST_K, CRB_INST_PART)(unsigned channels, int fmt, unsigned *block)
{
#if CRB_INST_PART == 0
	if (fmt == 2) return inst<0, 2>(block);             /* diagnostic format: channel count at run time */
	if (channels == 0 && CRB_INST_K <= 1) return fmt == 1 ? inst<0, 1>(block) : inst<0, 0>(block);
	switch (channels) {
	case 1: return fmt == 1 ? inst<1, 1>(block) : inst<1, 0>(block);
	case 2: return fmt == 1 ? inst<2, 1>(block) : inst<2, 0>(block);
	case 3: return fmt == 1 ? inst<3, 1>(block) : inst<3, 0>(block);
	case 4: return fmt == 1 ? inst<4, 1>(block) : inst<4, 0>(block);
	}
#elif CRB_INST_PART == 1
	if (fmt != 2)
		switch (channels) {
		case 5: return fmt == 1 ? inst<5, 1>(block) : inst<5, 0>(block);
		case 6: return fmt == 1 ? inst<6, 1>(block) : inst<6, 0>(block);
		case 7: return fmt == 1 ? inst<7, 1>(block) : inst<7, 0>(block);
		case 8: return fmt == 1 ? inst<8, 1>(block) : inst<8, 0>(block);
		}
#elif CRB_INST_PART == 2
	if (fmt != 2)
		switch (channels) {
		case 9: return fmt == 1 ? inst<9, 1>(block) : inst<9, 0>(block);
		case 10: return fmt == 1 ? inst<10, 1>(block) : inst<10, 0>(block);
		case 11: return fmt == 1 ? inst<11, 1>(block) : inst<11, 0>(block);
		case 12: return fmt == 1 ? inst<12, 1>(block) : inst<12, 0>(block);
		}
#else
	if (fmt != 2)
		switch (channels) {
		case 13: return fmt == 1 ? inst<13, 1>(block) : inst<13, 0>(block);
		case 14: return fmt == 1 ? inst<14, 1>(block) : inst<14, 0>(block);
		case 15: return fmt == 1 ? inst<15, 1>(block) : inst<15, 0>(block);
		case 16: return fmt == 1 ? inst<16, 1>(block) : inst<16, 0>(block);
		}
#endif
	return (crb_kernel_fn)NULL;
}
