// Exercises the C++ host mirror (include/hq_plugin.hpp) exactly as a C++ embedder would:
// hq::HybridQuantization::quantization() end to end, and the class-level API.  Prints one JSON
// object that tests/test_gpu_cpp_api.py compares with the oracle.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "hq_plugin.hpp"

static uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

int main(int argc, char** argv) {
    const int w = argc > 1 ? atoi(argv[1]) : 96, h = argc > 2 ? atoi(argv[2]) : 64;
    const int K = argc > 3 ? atoi(argv[3]) : 8, imax = argc > 4 ? atoi(argv[4]) : 120;
    // uniform synthetic image: same bytes as hybridquantization_b200.synth.synth_image(w, h, 5)
    std::vector<uint8_t> rgb((size_t)w * h * 3);
    uint64_t s = 5, cur = 0;
    splitmix64(s); splitmix64(s);
    for (size_t j = 0; j < rgb.size(); ++j) {
        if ((j & 7) == 0) cur = splitmix64(s);
        rgb[j] = (uint8_t)(cur >> (8 * (j & 7)));
    }
    try {
        hq::HybridQuantization plugin;
        plugin.nbOfColors = K;
        plugin.imax = imax;
        plugin.seed = 4242;
        std::vector<uint8_t> out(rgb.size());
        double bestError = 0;
        const std::vector<float> best = plugin.quantization(rgb.data(), w, h, out.data(), &bestError);
        unsigned long long sum = 0;
        for (size_t j = 0; j < out.size(); ++j) sum = sum * 1099511628211ULL + out[j];
        printf("{\"best_error\": \"%a\", \"image_hash\": %llu, \"best_colors\": [", bestError, sum);
        for (size_t i = 0; i < best.size(); ++i) printf("%s\"%a\"", i ? ", " : "", best[i]);
        printf("], ");
        {   // the same search fed with the plugin's own float planes (c/255, HybridQuantization.java:95-98)
            const size_t n = (size_t)w * h;
            std::vector<float> planes(3 * n);
            for (size_t j = 0; j < n; ++j)
                for (int c = 0; c < 3; ++c) planes[(size_t)c * n + j] = (float)(rgb[3 * j + c] / 255.0);
            std::vector<uint8_t> out2(rgb.size());
            double err2 = 0;
            const std::vector<float> best2 = plugin.quantization(planes.data(), planes.data() + n, planes.data() + 2 * n, w, h, out2.data(), &err2);
            printf("\"float_image_same\": %d, ", (int)(best2 == best && out2 == out && err2 == bestError));
        }
        {   // error-image mode (HybridQuantization.errorImage :139-182) between the input and its quantisation
            std::vector<uint8_t> map8((size_t)w * h);
            const double mean = plugin.errorImage(rgb.data(), out.data(), w, h, nullptr, map8.data());
            unsigned long long mh = 0;
            for (size_t j = 0; j < map8.size(); ++j) mh = mh * 1099511628211ULL + map8[j];
            printf("\"error_image_mean\": \"%a\", \"error_map_hash\": %llu, ", mean, mh);
        }
        // class-level API: SWASA draws + one population evaluation
        hq::JavaRandom rnd(77760);
        hq::SWASA swasa(4, 5000, 20, 2.0f, 0.75f, 0.15f, 20.0f, 0.9f, 100.0f, 5.3f, &rnd);
        std::vector<float> colors(4 * 4 * K);
        for (int i = 0; i < 4; ++i) swasa.generateRandomColors(K, colors.data() + (size_t)i * 4 * K);
        hq::ImageManipulation be(hq::ImageManipulation::deltaETypes::CIE76, false, true, 0);
        be.setImage(rgb.data(), w, h, HQ_WHITEPOINT_D65);
        const std::vector<double> costs = be.computeQuantizationErrorPopulation(4, colors.data(), K, swasa, (uint64_t)w * h, HQ_SPACE_LAB);
        printf("\"costs\": [\"%a\", \"%a\", \"%a\", \"%a\"]}\n", costs[0], costs[1], costs[2], costs[3]);
        // error behaviour: exceptions, not silent zeros
        try { be.quantize(colors.data(), 5000, HQ_SPACE_LAB); return 3; } catch (const std::runtime_error&) {}
    } catch (const std::exception& e) {
        fprintf(stderr, "FAILED: %s\n", e.what());
        return 2;
    }
    return 0;
}
