// Read-batching driver on the device side: conditioning -> 2 flank alignments per read ->
// count-HMM Viterbi -> optional methylation-HMM Viterbi (reference repeatCounter.detect,
// scripts/STRique.py:581-618, batched).
#pragma once
#include "align.cuh"
#include "condition.cuh"
#include "viterbi.cuh"

namespace strique {

struct Target {                       // one (locus, strand) classifier (S.py:561-575)
    std::vector<float> prefix_levels, suffix_levels;   // k-mer means of prefix_ext / suffix_ext
    int pre_trim = 0, post_trim = 0;  // len(prefix_ext) - len(prefix), len(suffix_ext) - len(suffix) in samples
    int count_model = -1, mod_model = -1;
    int count_offset = 0;             // flanking_count - repeat_offset (S.py:378, 437)
};

}  // namespace strique
