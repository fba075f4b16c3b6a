// hq_plugin.hpp — C++ host side of the plugin path, above the C ABI (hq_b200.h).
//
// The reference's host code is Java (no JDK in this image), so the host side is written
// in C++ and mirrors the reference's classes for this path — same names, same argument
// meaning, same draw order of the random generator:
//   hq::SWASA              <-> SWASA.java              (annealing schedule, neighbour moves)
//   hq::ImageManipulation  <-> ImageManipulation.java  (backend: findBestQuantization, quantize)
//   hq::ScielabProcessor   <-> ScielabProcessor.java   (white point, bestColors façade)
//   hq::HybridQuantization <-> HybridQuantization.java (the 20 plugin parameters + quantization())
//   hq::JavaRandom         <-> icy.util.Random (a static java.util.Random; un-vendored icy.jar)
// Header only; uses nothing but the C ABI.  Citations are File:line in the reference.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "hq_b200.h"

namespace hq {

// java.util.Random (JDK-specified 48-bit LCG).  The reference draws from the static,
// unseeded icy.util.Random (SWASA.java:46-48,56,61,96-98); an explicit seed is the one
// added parameter so that runs are reproducible.
class JavaRandom {
public:
    explicit JavaRandom(int64_t seed = 0) { setSeed(seed); }
    void setSeed(int64_t seed) { state_ = (static_cast<uint64_t>(seed) ^ 0x5DEECE66DULL) & kMask; }
    int32_t next(int bits) {
        state_ = (state_ * 0x5DEECE66DULL + 0xBULL) & kMask;
        return static_cast<int32_t>(static_cast<int64_t>(state_ >> (48 - bits)));
    }
    float nextFloat() { return static_cast<float>(next(24)) / static_cast<float>(1 << 24); }
    double nextDouble() {
        const int64_t hi = static_cast<int64_t>(next(26)) << 27;
        return static_cast<double>(hi + next(27)) * 0x1.0p-53;
    }
    uint64_t state() const { return state_; }
    void setState(uint64_t s) { state_ = s & kMask; }

private:
    static constexpr uint64_t kMask = (1ULL << 48) - 1;
    uint64_t state_ = 0;
};

// SWASA.java, member for member.
class SWASA {
public:
    SWASA(int population, int imax, int iTc, float delta, float convDelay, float convSpread,
          float t0, float alpha, float s0, float beta, JavaRandom* random)
        : population_(population), imax_(imax), iTc_(iTc), delta_(delta), t0_(t0), alpha_(alpha),
          s0_(s0), beta_(beta), convergenceDelay_(convDelay), convergenceRate_(convSpread),
          random_(random) {
        reset();
    }
    void reset() { temperature_ = t0_; }  // SWASA.java:30-34 (stepWidth is dead there)
    int getImax() const { return imax_; }
    int getPopulationSize() const { return population_; }
    float temperature() const { return temperature_; }

    // :40-52 — colors is [numberOfColors][4]
    void generateRandomColors(int numberOfColors, float* colors) {
        for (int i = 0; i < numberOfColors; ++i) {
            const int offset = i << 2;
            colors[offset] = random_->nextFloat();
            colors[offset + 1] = random_->nextFloat();
            colors[offset + 2] = random_->nextFloat();
            colors[offset + 3] = 0.0f;
        }
    }
    // :54-57 — draws from the generator only when deltaE > 0
    bool isAccepted(double deltaE) {
        return deltaE <= 0 || acceptanceProbability(deltaE) > random_->nextDouble();
    }
    // :59-62 — the tanh argument is formed in float, as Java's int/float promotion does
    bool keepsHisValues(int iteration) {
        const float num = static_cast<float>(iteration) - convergenceDelay_ * static_cast<float>(imax_);
        const float den = convergenceRate_ * static_cast<float>(imax_);
        return -(std::tanh(static_cast<double>(num / den))) / 2 + 0.5 > random_->nextDouble();
    }
    double acceptanceProbability(double deltaE) const {  // :64-67
        return std::exp(-deltaE / static_cast<double>(temperature_));
    }
    float maxStepWidth(int i) const {  // :69-72
        const float arg = beta_ * static_cast<float>(i) / static_cast<float>(imax_);
        const float twoS0 = 2.0f * s0_;
        return static_cast<float>(static_cast<double>(twoS0) / (1.0 + std::exp(static_cast<double>(arg))));
    }
    // :74-82 — the reference passes the int[] of used flags; a colour is used iff count > 0
    double computePenalty(const uint64_t* counts, int numberOfColors) const {
        double penalty = 0;
        for (int c = 0; c < numberOfColors; ++c)
            if (counts[c] == 0) penalty += delta_;
        return penalty;
    }
    void reduceTemperatureIfNecessary(int iteration) {  // :84-89
        if (iteration % iTc_ == 0) temperature_ *= alpha_;
    }
    // :91-101
    void generateNeighboringColors(const float* colors, float* nextColors, int numberOfColors, int iteration) {
        const float actualMaxStepWidth = maxStepWidth(iteration) / 256.0f;
        for (int i = 0; i < numberOfColors; ++i) {
            const int offset = i << 2;
            for (int ch = 0; ch < 3; ++ch) {
                const float u = random_->nextFloat() * 2.0f - 1.0f;
                nextColors[offset + ch] = clamp(colors[offset + ch] + u * actualMaxStepWidth, 0.0f, 1.0f);
            }
            nextColors[offset + 3] = 0.0f;
        }
    }
    static float clamp(float value, float min, float max) {  // :103-106
        return value > min ? (value > max ? max : value) : min;
    }
    float delta() const { return delta_; }

private:
    int population_, imax_, iTc_;
    float delta_, t0_, alpha_, s0_, beta_;
    float convergenceDelay_, convergenceRate_;
    float temperature_ = 0.f;
    JavaRandom* random_;
};

// ImageManipulation.java: the compute backend.  Owns (or borrows) one hq_ctx.
class ImageManipulation {
public:
    enum class deltaETypes { CIE76, CIE94, CIEDE2000 };  // :20 — the plugin only ever passes CIE76 (HybridQuantization.java:96)

    // :52 — creation failure THROWS; the reference's silent "pure Java mode" (:79-92) that
    // returns zero arrays does not exist here.
    ImageManipulation(deltaETypes deltaEType, bool verbose, bool convergence, int device = 0)
        : verbose_(verbose), convergence_(convergence), owns_(true) {
        if (hq_create(device, &ctx_) != HQ_OK) throw std::runtime_error(std::string("hq_create: ") + hq_last_error(nullptr));
        // :63 program.addBuildOption("-D" + deltaEType.name()); CIEDE2000 is an empty stub in the reference (cl:227-229) and refused here
        const int t = deltaEType == deltaETypes::CIE76 ? HQ_DELTAE_CIE76 : (deltaEType == deltaETypes::CIE94 ? HQ_DELTAE_CIE94 : HQ_DELTAE_CIEDE2000);
        if (hq_set_delta_e(ctx_, t) != HQ_OK) {
            const std::string msg = std::string("hq_set_delta_e: ") + hq_last_error(ctx_);
            hq_destroy(ctx_); ctx_ = nullptr;
            throw std::runtime_error(msg);
        }
    }
    // view over an existing context (used by the C ABI's hq_find_best_quantization)
    ImageManipulation(hq_ctx* borrowed, bool verbose, bool convergence)
        : ctx_(borrowed), verbose_(verbose), convergence_(convergence), owns_(false) {}
    ~ImageManipulation() { close(); }
    ImageManipulation(const ImageManipulation&) = delete;
    ImageManipulation& operator=(const ImageManipulation&) = delete;

    bool getCudaAvailable() const { return ctx_ != nullptr; }  // getOpenCLAvailable, :95
    hq_ctx* context() const { return ctx_; }
    void close() {  // :265-269
        if (owns_ && ctx_) hq_destroy(ctx_);
        ctx_ = nullptr;
    }
    void setStopFlag(const volatile bool* flag) { stopFlag_ = flag; }
    void setCostModel(int costModel) { costModel_ = costModel; }  // HQ_COST_LAB | HQ_COST_SCIELAB
    void setEvalFlags(int flags) { evalFlags_ = flags; }           // HQ_EVAL_* passed to hq_eval_palettes; -1 (default) = hq_search_eval_flags
    // updateProgressBar (HybridQuantization.java:265-270), invoked every 10 iterations (:546-551)
    void setProgress(void (*fn)(void*, int, int, double), void* user) { progress_ = fn; progressUser_ = user; }

    // upload + RGB->CIELAB (the uploads of :451,:471-472 and RGBtoXYZ/XYZtoScielab)
    void setImage(const uint8_t* rgb, int w, int rows, int whitepoint) {
        check(hq_set_image_u8(ctx_, rgb, w, rows, whitepoint), "hq_set_image_u8");
    }
    // the plugin's own representation: im.getDataXYCAsFloat() planes in [0,1] (HybridQuantization.java:95-98)
    void setImage(const float* r, const float* g, const float* b, int w, int rows, int whitepoint) {
        check(hq_set_image_f32_planar(ctx_, r, g, b, w, rows, whitepoint), "hq_set_image_f32_planar");
    }

    // :620-727 — evaluates the whole population in one launch
    std::vector<double> computeQuantizationErrorPopulation(int populationSize, const float* colors, int nbOfColors,
                                                           const SWASA& swasa, uint64_t nTotal, int space) {
        errFx_.resize(populationSize);
        counts_.resize(static_cast<size_t>(populationSize) * nbOfColors);
        if (costModel_ == HQ_COST_SCIELAB)  // the reference's own kernel chain (:635-699)
            check(hq_eval_palettes_scielab(ctx_, colors, populationSize, nbOfColors, space, errFx_.data(), counts_.data()), "hq_eval_palettes_scielab");
        else {
            const int flags = evalFlags_ >= 0 ? evalFlags_ : hq_search_eval_flags(ctx_, nbOfColors, space, costModel_);  // exact pruning where it pays
            check(hq_eval_palettes(ctx_, colors, populationSize, nbOfColors, space, flags, errFx_.data(), counts_.data(), nullptr), "hq_eval_palettes");
        }
        std::vector<double> results(populationSize);
        for (int i = 0; i < populationSize; ++i) {
            // :712 averageArray(err) + computePenalty(used)
            const uint64_t* cnt = counts_.data() + static_cast<size_t>(i) * nbOfColors;
            // (HQ_ERR_FX_NAN: a NaN pixel of the CIE94 branch — the reference's averageArray would return NaN)
            const double sum = errFx_[i] == HQ_ERR_FX_NAN ? std::nan("") : static_cast<double>(errFx_[i]) * (1.0 / 16777216.0);
            results[i] = sum / static_cast<double>(nTotal) + swasa.computePenalty(cnt, nbOfColors);
        }
        return results;
    }

    static int argmin(const std::vector<double>& arr) {  // :843-856
        int min = 0;
        double score = arr[0];
        for (size_t i = 1; i < arr.size(); ++i)
            if (score > arr[i]) { min = static_cast<int>(i); score = arr[i]; }
        return min;
    }

    // :383-591.  The image must have been set.  Returns the best palette [K][4].
    std::vector<float> findBestQuantization(int nbOfColors, SWASA& simulatedAnnealing, uint64_t nTotal, int space,
                                            double* bestErrorOut = nullptr, double* traceCosts = nullptr,
                                            int* iterationsDone = nullptr) {
        simulatedAnnealing.reset();  // :385
        const int populationSize = simulatedAnnealing.getPopulationSize();
        const size_t pal = static_cast<size_t>(nbOfColors) * 4;
        if (nTotal == 0) nTotal = hq_image_pixels(ctx_);
        std::vector<float> colors(populationSize * pal), currentColors(populationSize * pal), bestColors(pal);
        for (int i = 0; i < populationSize; ++i)  // :413-417
            simulatedAnnealing.generateRandomColors(nbOfColors, colors.data() + i * pal);
        std::vector<double> currentErrors =  // :490
            computeQuantizationErrorPopulation(populationSize, colors.data(), nbOfColors, simulatedAnnealing, nTotal, space);
        if (traceCosts) std::memcpy(traceCosts, currentErrors.data(), sizeof(double) * populationSize);
        const int mn = argmin(currentErrors);  // :491-493
        double bestError = currentErrors[mn];
        std::memcpy(bestColors.data(), colors.data() + mn * pal, sizeof(float) * pal);
        const int maxiter = simulatedAnnealing.getImax();
        int ite = 1;
        for (; ite <= maxiter; ++ite) {                      // :497
            if (stopFlag_ && *stopFlag_) break;              // :499-502
            simulatedAnnealing.reduceTemperatureIfNecessary(ite);  // :507
            for (int j = 0; j < populationSize; ++j)         // :508-511
                simulatedAnnealing.generateNeighboringColors(colors.data() + j * pal, currentColors.data() + j * pal, nbOfColors, ite);
            const std::vector<double> errors =               // :515
                computeQuantizationErrorPopulation(populationSize, currentColors.data(), nbOfColors, simulatedAnnealing, nTotal, space);
            if (traceCosts) std::memcpy(traceCosts + static_cast<size_t>(ite) * populationSize, errors.data(), sizeof(double) * populationSize);
            double minerror = DBL_MAX;  // :516
            int minerroridx = 0;
            for (int i = 0; i < populationSize; ++i) {       // :518-537
                if (populationSize > 1 && errors[i] < minerror) { minerror = errors[i]; minerroridx = i; }
                if (simulatedAnnealing.isAccepted(errors[i] - currentErrors[i])) {
                    currentErrors[i] = errors[i];
                    std::memcpy(colors.data() + i * pal, currentColors.data() + i * pal, sizeof(float) * pal);
                    if (currentErrors[i] < bestError) {
                        bestError = currentErrors[i];
                        std::memcpy(bestColors.data(), currentColors.data() + i * pal, sizeof(float) * pal);
                        if (verbose_) std::printf("Best Error :%.17g\n", bestError);
                    }
                }
            }
            for (int i = 0; convergence_ && populationSize > 1 && i < populationSize; ++i) {  // :538-545
                if (!simulatedAnnealing.keepsHisValues(ite)) {
                    currentErrors[i] = minerror;
                    std::memcpy(colors.data() + i * pal, currentColors.data() + minerroridx * pal, sizeof(float) * pal);
                }
            }
            if (ite % 10 == 0 && progress_) progress_(progressUser_, ite, maxiter, bestError);  // :546-551
        }
        if (verbose_) std::printf("Final error : %.5f\n", bestError);  // :589
        if (bestErrorOut) *bestErrorOut = bestError;
        if (iterationsDone) *iterationsDone = ite - 1;
        return bestColors;
    }

    // :770-798 — returns float RGBA per pixel like the reference; u8 / indices optional
    std::vector<float> quantize(const float* colors, int nbOfColors, int space, uint8_t* outRgb = nullptr,
                                uint16_t* outIdx = nullptr) {
        std::vector<float> quantizedImage(static_cast<size_t>(hq_image_pixels(ctx_)) * 4);
        check(hq_quantize(ctx_, colors, nbOfColors, space, outRgb, quantizedImage.data(), outIdx), "hq_quantize");
        return quantizedImage;
    }

    // updateOpenCLFilters (:800-841): the 7 separable filters [7][taps] and |k3| [taps] of the S-CIELAB stage
    void updateFilters(const float* filters7, const float* absfilters, int taps) {
        check(hq_scielab_set_filters(ctx_, filters7, absfilters, taps), "hq_scielab_set_filters");
    }
    // XYZtoScielab(RGBtoXYZ(.)) of the resident image (:100-153, :285-370): planes [3][pixels]
    std::vector<float> scielabImage() {
        std::vector<float> planes(static_cast<size_t>(hq_image_pixels(ctx_)) * 3);
        check(hq_scielab_get_image(ctx_, planes.data()), "hq_scielab_get_image");
        return planes;
    }
    // computeError (:858-894): mean dE between S-CIELAB(resident image) and S-CIELAB(other image of the same size);
    // errorMap / errorMapU8 (optional, one value per pixel) receive ((255 - dE)^2) / 255^2 of :890
    double computeError(const uint8_t* otherRgb, float* errorMap = nullptr, uint8_t* errorMapU8 = nullptr) {
        double mean = 0;
        check(hq_error_image(ctx_, otherRgb, errorMap, errorMapU8, &mean), "hq_error_image");
        return mean;
    }
    double computeError(const float* r, const float* g, const float* b, float* errorMap = nullptr, uint8_t* errorMapU8 = nullptr) {
        double mean = 0;
        check(hq_error_image_f32_planar(ctx_, r, g, b, errorMap, errorMapU8, &mean), "hq_error_image_f32_planar");
        return mean;
    }

private:
    void check(int rc, const char* what) const {
        if (rc != HQ_OK) throw std::runtime_error(std::string(what) + ": " + hq_last_error(ctx_));
    }
    hq_ctx* ctx_ = nullptr;
    bool verbose_, convergence_, owns_;
    const volatile bool* stopFlag_ = nullptr;
    int costModel_ = HQ_COST_LAB;
    int evalFlags_ = -1;
    void (*progress_)(void*, int, int, double) = nullptr;
    void* progressUser_ = nullptr;
    std::vector<int64_t> errFx_;
    std::vector<uint64_t> counts_;
};

// ScielabProcessor.java: white point (:19-21,70-76), the S-CIELAB separable filter bank
// (:78-181, helpers :185-254) and the bestColors façade (:383-386).
class ScielabProcessor {
public:
    enum class Whitepoint { D50, D65 };
    // The filter bank: Ofilters[channel][gaussian][tap] and |Ofilters[0][2]| (:54-55)
    struct FilterBank {
        std::vector<float> Ofilters[3][3];
        std::vector<float> absOfilters;
        int taps() const { return static_cast<int>(Ofilters[0][0].size()); }
        // flattened as the C ABI takes it: [7][taps] = O1g1,O1g2,O1g3,O2g1,O2g2,O3g1,O3g2
        std::vector<float> flat() const {
            std::vector<float> f;
            const int count[3] = {3, 2, 2};
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < count[i]; ++j) f.insert(f.end(), Ofilters[i][j].begin(), Ofilters[i][j].end());
            return f;
        }
    };

    ScielabProcessor(int dpi, double viewingDistance, Whitepoint whitepoint, ImageManipulation* imageProcessor)
        : dpi_(dpi), viewingDistance_(viewingDistance), whitepoint_(whitepoint), imageProcessing_(imageProcessor),
          bank_(buildFilters(dpi, viewingDistance)) {}

    // :238-254 — a centred Gaussian that sums to one
    static std::vector<float> gauss(float halfwidth, int width) {
        const float alpha = 2 * static_cast<float>(std::sqrt(std::log(2.0))) / (halfwidth - 1);
        std::vector<float> result(width);
        const int offset = width / 2;
        double sum = 0;
        for (int i = 0; i < width; ++i) {
            const float e = -alpha * alpha * static_cast<float>(i - offset) * static_cast<float>(i - offset);
            result[i] = static_cast<float>(std::exp(static_cast<double>(e)));
            sum += result[i];
        }
        for (float& v : result) v = static_cast<float>(static_cast<double>(v) / sum);
        return result;
    }

    // :66-181 — everything the constructor computes before handing the filters to the backend
    static FilterBank buildFilters(int dpi, double viewingDistance) {
        static const float weights[3][3] = {{1.00327f, 0.114416f, -0.117686f}, {0.616725f, 0.383275f, 0}, {0.567885f, 0.432115f, 0}};  // :44-48
        static const float halfwidths[3][3] = {{0.05f, 0.225f, 7.0f}, {0.0685f, 0.826f, 0}, {0.0920f, 0.6451f, 0}};                      // :49-53
        static const int count[3] = {3, 2, 2};
        const double kPi = 3.14159265358979323846;
        int sampPerDeg = static_cast<int>(std::llround(dpi / ((180 / kPi) * std::atan(2.54 / viewingDistance))));  // :80
        int uprate = 1;
        if (sampPerDeg < 224) {  // :81-88 (minSAMPPERDEG :23)
            uprate = static_cast<int>(std::ceil(224 * 1.0 / sampPerDeg));
            sampPerDeg *= uprate;
        }
        const int width = static_cast<int>(std::ceil(sampPerDeg / 2.0)) * 2 - 1;  // :102
        FilterBank bank;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < count[i]; ++j) {
                std::vector<float> f = gauss(halfwidths[i][j] * static_cast<float>(sampPerDeg), width);  // :97,112
                const float w = weights[i][j];
                const float sign = w > 0 ? 1.0f : (w < 0 ? -1.0f : 0.0f);
                const float factor = static_cast<float>(std::sqrt(static_cast<double>(std::fabs(w)))) * sign;  // :113
                for (float& v : f) v *= factor;
                bank.Ofilters[i][j] = std::move(f);
            }
        if (uprate > 1) {  // :122-173: triangular up-sampling kernel, convolve, keep every uprate-th sample around the centre
            const int upLen = uprate * 2 - 1;
            std::vector<float> upcol(upLen);
            for (int i = 0; i < upLen; ++i) upcol[i] = static_cast<float>(uprate - std::abs(uprate - i - 1)) * 1.0f / static_cast<float>(uprate);  // :129
            upcol = resize1D(upcol, upLen + width - 1);  // :132
            const int mid = width / 2;
            std::vector<int> downs;  // :146-164
            for (int v = mid - (mid / uprate) * uprate; v <= mid; v += uprate) downs.push_back(v);
            for (int v = mid + uprate; static_cast<int>(downs.size()) < 2 * (mid / uprate) + 1; v += uprate) downs.push_back(v);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < count[i]; ++j) {
                    const std::vector<float> up = conv1D(bank.Ofilters[i][j], upcol);  // :136-144
                    std::vector<float> picked(downs.size());
                    for (size_t q = 0; q < downs.size(); ++q) picked[q] = up[downs[q]];  // :166-172, extractWithIndices :222-230
                    bank.Ofilters[i][j] = std::move(picked);
                }
        }
        bank.absOfilters.resize(bank.Ofilters[0][2].size());  // :174-178
        for (size_t i = 0; i < bank.absOfilters.size(); ++i) {
            const float v = bank.Ofilters[0][2][i];
            bank.absOfilters[i] = v * (v < 0 ? -1 : 1);
        }
        return bank;
    }

    const FilterBank& filters() const { return bank_; }
    int whitepointCode() const { return whitepoint_ == Whitepoint::D50 ? HQ_WHITEPOINT_D50 : HQ_WHITEPOINT_D65; }
    // sRGBToScielab (:374-381) with the identity filter: uploads the image, converts on the GPU
    void sRGBToScielab(const uint8_t* rgb, int w, int rows) { imageProcessing_->setImage(rgb, w, rows, whitepointCode()); }
    // float[][] sRGBImage as the reference passes it (:374): one plane per channel
    void sRGBToScielab(const float* r, const float* g, const float* b, int w, int rows) { imageProcessing_->setImage(r, g, b, w, rows, whitepointCode()); }
    std::vector<float> bestColors(int nbOfColors, SWASA& simulatedAnnealing, uint64_t nTotal, int space, double* bestError = nullptr) {
        return imageProcessing_->findBestQuantization(nbOfColors, simulatedAnnealing, nTotal, space, bestError);
    }
    void close() { imageProcessing_->close(); }  // :440-443

private:
    static std::vector<float> conv1D(const std::vector<float>& data, const std::vector<float>& filter) {  // :185-201
        std::vector<float> result(data.size(), 0.0f);
        const int n = static_cast<int>(data.size()), offset = static_cast<int>(filter.size()) / 2;
        for (int i = 0; i < n; ++i)
            for (int j = -offset; j <= offset; ++j)
                if (i + j >= 0 && i + j < n) result[i] += filter[j + offset] * data[i + j];
        return result;
    }
    static std::vector<float> resize1D(const std::vector<float>& src, int newSize) {  // :203-220
        std::vector<float> res(newSize, 0.0f);
        const int n = static_cast<int>(src.size());
        const int pad = std::abs(newSize - n) / 2;
        if (newSize > n) {
            for (int i = 0; i < n; ++i) res[pad + i] = src[i];
        } else {
            for (int i = 0; i < newSize; ++i) res[i] = src[pad + i];
        }
        return res;
    }
    int dpi_;
    double viewingDistance_;
    Whitepoint whitepoint_;
    ImageManipulation* imageProcessing_;
    FilterBank bank_;
};

// HybridQuantization.java: the parameter surface (:185-257, unchanged names, defaults and
// ranges) and quantization() (:93-137) without the Icy GUI objects.
struct HybridQuantization {
    int nbOfColors = 8;          // "Number of colors" [1, 2^24]  :192
    int populationSize = 4;      // "Population size"             :197
    int imax = 5000;             // "Max iterations"              :199
    float delta = 2.0f;          // "Penalty Constant"            :201
    bool convEnable = true;      // "Pop Convergence"             :204
    float convDelay = 0.75f;     // "Convergence delay"           :206
    float convSpread = 0.15f;    // "Convergence spread"          :208
    float T0 = 20.0f;            // "Initial temperature"         :212
    int iTc = 20;                // "Iterations per temperature"  :214
    float alpha = 0.9f;          // "Cooling coefficient"         :216
    float s0 = 100.0f;           // "Initial Step size"           :223
    float beta = 5.3f;           // "Adaptation constant"         :224
    int dpi = 72;                // "Dpi"                         :229
    float viewingDistance = 45;  // "Viewing distance"            :231
    ScielabProcessor::Whitepoint whitePoint = ScielabProcessor::Whitepoint::D65;  // :233
    bool verbose = false;        // "Verbosity"                   :237
    // added (not in the reference): reproducibility + assignment space + device
    int64_t seed = 77760;
    int space = HQ_SPACE_LAB;
    int costModel = HQ_COST_LAB;  // HQ_COST_SCIELAB = score exactly like the reference plugin (use with HQ_SPACE_SRGB)
    int device = 0;
    volatile bool stopFlag = false;  // :52

    void stopExecution() { stopFlag = true; }  // :311-315
    bool isStopFlag() const { return stopFlag; }

    // quantization(), :93-137: image in (packed u8 RGB), quantised image out (packed u8 RGB)
    // plus the best palette.  Throws std::invalid_argument on the reference's input checks
    // (:65-70: an image with at least 3 channels is required).
    std::vector<float> quantization(const uint8_t* rgb, int w, int h, uint8_t* outRgb, double* bestError = nullptr) {
        if (!rgb || w <= 0 || h <= 0) throw std::invalid_argument("Please open an image first.");
        return run(w, h, outRgb, bestError, [&](ScielabProcessor& sp) { sp.sRGBToScielab(rgb, w, h); });
    }
    // the same on the converted float image the plugin holds (im.getDataXYCAsFloat(), :95-98): one plane per channel
    std::vector<float> quantization(const float* r, const float* g, const float* b, int w, int h, uint8_t* outRgb, double* bestError = nullptr) {
        if (!r || !g || !b || w <= 0 || h <= 0) throw std::invalid_argument("Please open an image first.");
        return run(w, h, outRgb, bestError, [&](ScielabProcessor& sp) { sp.sRGBToScielab(r, g, b, w, h); });
    }

    // errorImage(), :139-182: mean S-CIELAB dE between two images of the same size and the error map of :890
    double errorImage(const uint8_t* original, const uint8_t* quantized, int w, int h, float* errorMap = nullptr, uint8_t* errorMapU8 = nullptr) {
        if (!original) throw std::invalid_argument("Please open/select the original image first.");    // :76-77
        if (!quantized) throw std::invalid_argument("Please open/select the quantized image first.");   // :78-79
        if (w <= 0 || h <= 0) throw std::invalid_argument("Mismatching image sizes or not enough channels, abort.");  // :80-81
        ImageManipulation imageProcessor(ImageManipulation::deltaETypes::CIE76, verbose, false, device);  // :145
        ScielabProcessor scielabProcessor(dpi, viewingDistance, whitePoint, &imageProcessor);              // :146
        scielabProcessor.sRGBToScielab(original, w, h);                                                    // :148
        const std::vector<float> flat = scielabProcessor.filters().flat();
        imageProcessor.updateFilters(flat.data(), scielabProcessor.filters().absOfilters.data(), scielabProcessor.filters().taps());
        const double mean = imageProcessor.computeError(quantized, errorMap, errorMapU8);                  // :151-160
        scielabProcessor.close();
        return mean;
    }

  private:
    template <class Upload>
    std::vector<float> run(int w, int h, uint8_t* outRgb, double* bestError, Upload upload) {
        (void)w; (void)h;
        stopFlag = false;
        ImageManipulation imageProcessor(ImageManipulation::deltaETypes::CIE76, verbose, convEnable, device);  // :96
        imageProcessor.setStopFlag(&stopFlag);
        imageProcessor.setCostModel(costModel);
        JavaRandom random(seed);
        SWASA swasa(populationSize, imax, iTc, delta, convDelay, convSpread, T0, alpha, s0, beta, &random);  // :97
        ScielabProcessor scielabProcessor(dpi, viewingDistance, whitePoint, &imageProcessor);               // :101
        upload(scielabProcessor);                                                                           // :104
        if (costModel == HQ_COST_SCIELAB) {  // :180 updateOpenCLFilters with the bank built from dpi / viewing distance
            const std::vector<float> flat = scielabProcessor.filters().flat();
            if (hq_scielab_set_filters(imageProcessor.context(), flat.data(), scielabProcessor.filters().absOfilters.data(), scielabProcessor.filters().taps()) != HQ_OK)
                throw std::runtime_error(std::string("hq_scielab_set_filters: ") + hq_last_error(imageProcessor.context()));
        }
        std::vector<float> best = scielabProcessor.bestColors(nbOfColors, swasa, 0, space, bestError);     // :107
        if (outRgb) imageProcessor.quantize(best.data(), nbOfColors, space, outRgb);                        // :109
        scielabProcessor.close();                                                                           // :136
        return best;
    }
};

}  // namespace hq
