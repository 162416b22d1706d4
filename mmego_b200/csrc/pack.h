// pack.h -- host-side weight packing (pure C++; unit-tested on CPU through the mmego_pack_* test hooks).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "internal.h"

namespace mmego {

struct StateDict {
    HostSD m;
    const float* get(const std::string& name, long long numel) const {
        auto it = m.find(name);
        if (it == m.end()) throw std::runtime_error("missing state_dict tensor: " + name);
        if (it->second.second != numel)
            throw std::runtime_error("state_dict tensor " + name + " has " + std::to_string(it->second.second) +
                                     " elements, expected " + std::to_string(numel));
        return it->second.first;
    }
};

struct BnAffine {
    std::vector<float> s, o;
};
BnAffine bn_affine(const StateDict& sd, const std::string& prefix, int C);

// W [N][sum(ksegs)] row-major -> segments padded to 16; optional per-row scale and bias transform
HostPackedGemm pack_linear(const float* W, const float* b, int N, const std::vector<int>& ksegs,
                           const float* row_scale = nullptr, const float* row_offset = nullptr);

struct HostBigLstm {
    HostPackedGemm dir[2];
};
HostBigLstm pack_big_lstm(const StateDict& sd, const std::string& prefix, int layer, int In, int H);

struct HostSmallLstm {
    HostPackedGemm ih;
    std::vector<float> whh;
    int in = 0;
};
HostSmallLstm pack_small_lstm(const StateDict& sd, const std::string& prefix, int layer, int In);

std::vector<float> pack_small_lstm_mma(const StateDict& sd, const std::string& prefix, int layer, int In);
std::vector<float> pack_imu_fc1_mma(const HostPackedGemm& fc1);
struct HeadLayerSpec {
    const float* W;   // [N][K] row-major (torch nn.Linear.weight)
    const float* b;   // [N]
    int N, K;
};
std::vector<float> pack_head_mma(const std::vector<HeadLayerSpec>& layers);
std::vector<float> pack_upper_point(const StateDict& sd);
std::vector<float> pack_lower_frame(const StateDict& sd);
// mma.sync (fp16 hi/lo fragment) packing of the folded blobs above (point_layout.h: UpperMmaLayout / LowerMmaLayout)
float pack_mma_weight(const float* W, int ldw, const std::vector<int>& kmap, const std::vector<int>& nmap, float* out);
std::vector<float> pack_upper_point_mma(const std::vector<float>& folded);
std::vector<float> pack_lower_frame_mma(const std::vector<float>& folded);

struct HostGcnLayer {
    std::vector<float> ahat;
    HostPackedGemm gconv, tconv;
    int cin = 0, cout = 0;
};
HostGcnLayer pack_gcn_layer(const StateDict& sd, const std::string& gcn_prefix, int layer, int Cin, int Cout);
std::vector<float> pack_data_bn(const StateDict& sd, const std::string& gcn_prefix);

}  // namespace mmego
