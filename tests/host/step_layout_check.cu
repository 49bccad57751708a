// Host check of the fused step kernel's shared-memory layout (omg_planner_b200/csrc/chomp_kernels.cuh: make_layout):
// prints the footprint of the CTA shapes launch_step picks for the BASELINE configurations and verifies that the
// regions do not overlap and that the scratch behind the two block_sum buffers holds the info row.
#include <cstdio>

#include "../../omg_planner_b200/csrc/chomp_kernels.cuh"

using namespace omgb;

static int check(int n, int c, int nobj, int nwarps, bool topk, int ctas) {
    const SmemLayout L = make_layout(n, c, 16, nobj, 16, nwarps, topk, false);
    int bad = 0;
    // regions in layout order (grad / u / viol alias the frames region by design)
    const unsigned seq[] = {L.off_xi, L.off_start, L.off_end, L.off_goal, L.off_frames, L.off_lg, L.off_red, L.off_mask,
                            L.off_best, L.off_act, L.off_win, L.off_bestp, L.off_objs, L.off_sph, L.off_hist, L.off_mbar,
                            L.total};
    for (unsigned k = 0; k + 1 < sizeof(seq) / sizeof(seq[0]); ++k) bad += seq[k] > seq[k + 1];
    bad += (L.off_viol + sizeof(double) * 3 * n * ND > L.off_lg) ? 1 : 0;              // viol + 2 x scan scratch inside frames
    bad += (L.off_mask - L.off_red < sizeof(double) * (L.red_max + 16)) ? 1 : 0;        // sum buffers + info row
    bad += (L.red_max != 2 * nwarps * 8) ? 1 : 0;
    bad += (L.off_lg % 16 || L.off_objs % 16 || L.off_sph % 16 || L.off_mbar % 8 || L.off_frames % 8) ? 1 : 0;
    const unsigned need = (unsigned)ctas * (L.total + 1024u);
    printf("n %d c %d objects %d warps %d topk %d total %u ctas %d need %u fits %d bad %d\n", n, c, nobj, nwarps, (int)topk,
           L.total, ctas, need, need <= 228u * 1024u ? 1 : 0, bad);
    return bad;
}

int main() {
    int bad = 0;
    bad += check(30, 5, 10, 10, true, 3);    // config 2: 320 threads x 3 CTAs/SM
    bad += check(60, 5, 20, 16, true, 2);    // config 4: 512 threads x 2
    bad += check(50, 5, 30, 16, true, 2);    // config 5
    bad += check(30, 0, 10, 10, false, 3);   // full-sum mode, fixed goal
    bad += check(30, 5, 10, 32, true, 1);    // small batches: 1024 threads x 1
    bad += check(2, 1, 1, 32, true, 1);
    bad += check(60, 5, 40, 16, true, 2);    // two mask words
    return bad ? 1 : 0;
}
