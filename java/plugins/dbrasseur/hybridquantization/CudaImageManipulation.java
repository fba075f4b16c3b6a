package plugins.dbrasseur.hybridquantization;

/**
 * Drop-in CUDA backend for the plugin: the public surface of {@code ImageManipulation} (ImageManipulation.java) — same
 * constructor, same method names, same argument lists, same array layouts — bound to libhq_b200.so through the JNI shim
 * java/jni/hq_jni.c.  HybridQuantization.java and ScielabProcessor.java need ONE change each: the type they instantiate /
 * hold (INTEGRATION.md).
 *
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK).  The JNI shim underneath it is compiled and every native below is
 * executed on the GPU through a fake JNIEnv (tests/cpp/jni_harness.c); this file is the thin Java face of those natives.
 *
 * The annealing loop, the RNG (icy.util.Random inside the UNCHANGED SWASA.java) and the accept/reject logic stay in Java,
 * as the north star asks: only computeQuantizationErrorPopulation (ImageManipulation.java:620-727) changes — ONE native
 * call scores the whole population on the GPU(s).  There is no "pure Java mode": where the reference swallows a failed
 * OpenCL initialisation and later returns zero arrays (:79-92, :397, :773), this class throws.
 */
public class CudaImageManipulation implements AutoCloseable {
    static { System.loadLibrary("hq_jni"); }

    public enum deltaETypes {CIE76, CIE94, CIEDE2000}

    /** what a candidate is scored with; SCIELAB is what the reference computes, LAB the identity-filter cost of the accelerated path */
    public enum CostModel {LAB, SCIELAB}

    /** receives (iteration, maxIterations, bestError) every 10 iterations of the library-side search */
    public interface ProgressListener { void progress(int iteration, int maxIterations, double bestError); }

    private static final int SPACE_LAB = 0, SPACE_SRGB = 1;

    private long ctx; // hq_ctx*
    private final boolean verbose;
    private final boolean convergence;
    private boolean filtersReady;
    private CostModel costModel = CostModel.SCIELAB;
    private int whitepoint = 0; // HQ_WHITEPOINT_D65; set from the illuminant findBestQuantization receives

    /** the reference's constructor (ImageManipulation.java:52): devices from -Dhq.devices=0,1,... (default: device 0) */
    public CudaImageManipulation(deltaETypes deltaEType, boolean verbose, boolean convergence) {
        this(deltaEType, verbose, convergence, devicesFromProperty());
    }

    /** several devices = ONE context: rows are split and the integer partials all-reduced (NCCL) inside the library */
    public CudaImageManipulation(deltaETypes deltaEType, boolean verbose, boolean convergence, int[] devices) {
        this.verbose = verbose;
        this.convergence = convergence;
        ctx = nCreate(devices); // throws RuntimeException when no usable device exists
        // program.addBuildOption("-D" + deltaEType.name()) (:63).  CIE94 is the reference kernel's branch bit for bit, latent NaN
        // included; CIEDE2000 is an empty stub in the reference (cl:227-229) and is refused (RuntimeException)
        try { nSetDeltaE(ctx, deltaEType.ordinal()); } catch (RuntimeException e) { close(); throw e; }
    }

    private static int[] devicesFromProperty() {
        String p = System.getProperty("hq.devices", "0");
        String[] parts = p.split(",");
        int[] d = new int[parts.length];
        for (int i = 0; i < parts.length; i++) d[i] = Integer.parseInt(parts[i].trim());
        return d;
    }

    /** ImageManipulation.getOpenCLAvailable (:95): a live context exists (construction would have thrown otherwise) */
    boolean getOpenCLAvailable() { return ctx != 0; }

    public int deviceCount() { return nDeviceCount(ctx); }

    public void setCostModel(CostModel m) { costModel = m; }

    /** EzStoppable.stopExecution for the library-side search; the Java-side loop polls the plugin's flag as the reference does */
    public void requestStop() { nRequestStop(ctx); }

    // ------------------------------------------------------------------ one-shot conversions (reference signatures)

    /** ImageManipulation.RGBtoXYZ (:100-152) */
    public float[] RGBtoXYZ(float[] R, float[] G, float[] B) {
        float[] output = new float[4 * R.length];
        nRgbToXyz(ctx, R, G, B, output);
        return output;
    }

    /** ImageManipulation.XYZtoScielab (:285-370) */
    public float[] XYZtoScielab(float[] XYZ, float[][][] filters, float[] absfilters, int w, float[] illuminant) {
        if (!filtersReady) updateOpenCLFilters(filters, absfilters);
        float[] lab = new float[XYZ.length];
        nXyzToScielab(ctx, XYZ, w, illuminant, lab);
        return lab;
    }

    /** ImageManipulation.updateOpenCLFilters (:800-841): filters = {{O1g1,O1g2,O1g3},{O2g1,O2g2},{O3g1,O3g2}}, absfilters = |O1g3| */
    public void updateOpenCLFilters(float[][][] filters, float[] absfilters) {
        final int taps = absfilters.length;
        float[] flat = new float[7 * taps];
        float[][] order = {filters[0][0], filters[0][1], filters[0][2], filters[1][0], filters[1][1], filters[2][0], filters[2][1]};
        for (int i = 0; i < 7; i++) System.arraycopy(order[i], 0, flat, i * taps, taps);
        nScielabSetFilters(ctx, flat, absfilters, taps);
        filtersReady = true;
    }

    // ------------------------------------------------------------------ the search (reference signature)

    /**
     * ImageManipulation.findBestQuantization (:383-591).  inlinergbOriginal: RGBA-interleaved floats in [0,1]
     * (HybridQuantization.makeinline, :279-291); inlineScielabOriginal: the S-CIELAB image XYZtoScielab returned.
     * Returns the best palette, RGBA-interleaved (float[4 * nbOfColors]).
     */
    public float[] findBestQuantization(float[] inlinergbOriginal, float[] inlineScielabOriginal, int w, int nbOfColors,
                                        SWASA simulatedAnnealing, float[][][] filters, float[] absfilters, float[] illuminant) {
        simulatedAnnealing.reset();
        if (!filtersReady) updateOpenCLFilters(filters, absfilters);
        upload(inlinergbOriginal, w, illuminant);
        if (costModel == CostModel.SCIELAB) nScielabSetImage(ctx, inlineScielabOriginal);
        final int population = simulatedAnnealing.getPopulationSize();
        final int stride = 4 * nbOfColors;
        float[][] colors = new float[population][];
        float[][] candidates = new float[population][stride];
        for (int j = 0; j < population; j++) colors[j] = simulatedAnnealing.generateRandomColors(nbOfColors);
        double[] currentErrors = computeQuantizationErrorPopulation(colors, nbOfColors, simulatedAnnealing);
        int first = 0;
        for (int j = 1; j < population; j++) if (currentErrors[first] > currentErrors[j]) first = j; // argmin, :843-856
        double bestError = currentErrors[first];
        float[] bestColors = colors[first].clone();
        final int maxiter = simulatedAnnealing.getImax();
        final long start = System.currentTimeMillis();
        for (int ite = 1; ite <= maxiter; ite++) {
            if (simulatedAnnealing.getPlugin().isStopFlag()) break; // :499
            simulatedAnnealing.reduceTemperatureIfNecessary(ite);
            for (int j = 0; j < population; j++) simulatedAnnealing.generateNeighboringColors(colors[j], candidates[j], nbOfColors, ite);
            double[] errors = computeQuantizationErrorPopulation(candidates, nbOfColors, simulatedAnnealing); // ONE native call
            double minerror = Double.MAX_VALUE;
            int minerroridx = 0;
            for (int i = 0; i < population; i++) {
                if (population > 1 && errors[i] < minerror) { minerror = errors[i]; minerroridx = i; }
                if (simulatedAnnealing.isAccepted(errors[i] - currentErrors[i])) {
                    currentErrors[i] = errors[i];
                    System.arraycopy(candidates[i], 0, colors[i], 0, stride);
                    if (currentErrors[i] < bestError) {
                        bestError = currentErrors[i];
                        System.arraycopy(candidates[i], 0, bestColors, 0, stride);
                        if (verbose) System.out.println("Best Error :" + bestError);
                    }
                }
            }
            for (int i = 0; convergence && population > 1 && i < population; i++) {
                if (!simulatedAnnealing.keepsHisValues(ite)) {
                    currentErrors[i] = minerror;
                    System.arraycopy(candidates[minerroridx], 0, colors[i], 0, stride);
                }
            }
            if (ite % 10 == 0) { // :546-551
                long left = (long) ((((System.currentTimeMillis() - start) * 1.0) / ite) * (maxiter - ite));
                String eta = (left / 60000 > 0 ? left / 60000 + "m" : "") + ((left % 60000) / 1000 + "s") + " restant";
                simulatedAnnealing.getPlugin().updateProgressBar(ite + "/" + maxiter + " : " + eta, (ite * 1.0) / maxiter);
            }
        }
        System.out.println("Final error : " + bestError);
        return bestColors;
    }

    /**
     * The same search run entirely inside the library (hq_find_best_quantization: SWASA's schedule with java.util.Random and
     * an explicit seed — reproducible, which the reference's unseeded static generator is not).  params: the plugin's values.
     */
    public float[] findBestQuantizationSeeded(float[] inlinergbOriginal, float[] inlineScielabOriginal, int w, int nbOfColors, int population,
                                              int imax, int iTc, float delta, float convDelay, float convSpread, float T0, float alpha, float s0,
                                              float beta, float[] illuminant, long seed, ProgressListener listener) {
        upload(inlinergbOriginal, w, illuminant);
        final boolean sc = costModel == CostModel.SCIELAB;
        if (sc && inlineScielabOriginal != null) nScielabSetImage(ctx, inlineScielabOriginal);
        int[] ip = {population, imax, iTc, convergence ? 1 : 0, sc ? SPACE_SRGB : SPACE_LAB, sc ? 1 : 0};
        float[] fp = {delta, convDelay, convSpread, T0, alpha, s0, beta};
        float[] best = new float[4 * nbOfColors];
        nFindBestQuantization(ctx, nbOfColors, ip, fp, seed, best, new double[1], null, listener);
        return best;
    }

    /** computeQuantizationErrorPopulation (:620-727): averageArray(err) + computePenalty(used) per candidate (:712) */
    private double[] computeQuantizationErrorPopulation(float[][] colors, int nbOfColors, SWASA simulatedAnnealing) {
        final int p = colors.length;
        float[] flat = new float[p * 4 * nbOfColors];
        for (int i = 0; i < p; i++) System.arraycopy(colors[i], 0, flat, i * 4 * nbOfColors, 4 * nbOfColors);
        long[] errFx = new long[p];
        long[] counts = new long[p * nbOfColors];
        final boolean sc = costModel == CostModel.SCIELAB;
        nEvalPalettes(ctx, flat, p, nbOfColors, sc ? SPACE_SRGB : SPACE_LAB, sc ? 1 : 0, errFx, counts);
        double[] results = new double[p];
        final long n = nPixels(ctx);
        int[] used = new int[nbOfColors];
        for (int i = 0; i < p; i++) {
            for (int k = 0; k < nbOfColors; k++) used[k] = counts[i * nbOfColors + k] != 0 ? 1 : 0;
            // Long.MIN_VALUE = HQ_ERR_FX_NAN: a NaN pixel of the CIE94 branch — the reference's averageArray would return NaN
            final double sum = errFx[i] == Long.MIN_VALUE ? Double.NaN : errFx[i] * (1.0 / 16777216.0);
            results[i] = sum / n + simulatedAnnealing.computePenalty(used); // SWASA.java:74-82
        }
        return results;
    }

    // ------------------------------------------------------------------ output image and error image (reference signatures)

    /** ImageManipulation.quantize (:770-798): each pixel <- its nearest palette colour by sRGB distance, RGBA-interleaved floats */
    public float[] quantize(float[] inlineImageRGB, float[] colors) {
        // the reference takes no width here: pixels are independent, so any factorisation of n serves the upload
        upload(inlineImageRGB, inlineImageRGB.length / 4, null);
        float[] quantizedImage = new float[inlineImageRGB.length];
        nQuantize(ctx, colors, colors.length / 4, SPACE_SRGB, null, quantizedImage);
        return quantizedImage;
    }

    /** ImageManipulation.computeError (:858-894): mean CIE76 between two Lab images, error image in lanes 0..2 */
    public double computeError(float[] original, float[] quantized, float[] errorImage) {
        return nDeltaEImages(ctx, original, quantized, errorImage);
    }

    /** ImageManipulation.close (:265-269) */
    @Override public void close() { if (ctx != 0) { nDestroy(ctx); ctx = 0; } }

    // ------------------------------------------------------------------ helpers
    private float[] lastUploaded; // the very array object last uploaded: bestColors() and quantize() pass the same one (HybridQuantization.java:107-109)

    /** RGBA-interleaved image -> planes -> device (replaces the uploads at :451, :471-472) */
    private void upload(float[] inlineRGB, int w, float[] illuminant) {
        if (illuminant != null) whitepoint = illuminant[2] < 1.0f ? 1 : 0; // D50 Z = 0.825188, D65 Z = 1.0883 (ScielabProcessor.java:20-21)
        if (inlineRGB == lastUploaded) return;
        final int n = inlineRGB.length / 4;
        float[] r = new float[n], g = new float[n], b = new float[n];
        for (int i = 0, off = 0; i < n; i++, off += 4) { r[i] = inlineRGB[off]; g[i] = inlineRGB[off + 1]; b[i] = inlineRGB[off + 2]; }
        nSetImageFloat(ctx, r, g, b, w, n / w, whitepoint);
        lastUploaded = inlineRGB;
    }

    private static native long nCreate(int[] devices);
    private static native void nDestroy(long ctx);
    private static native long nPixels(long ctx);
    private static native int nDeviceCount(long ctx);
    private static native void nSetPruning(long ctx, int mode);
    private static native void nSetDeltaE(long ctx, int type);
    private static native void nRequestStop(long ctx);
    private static native void nSetImage(long ctx, byte[] rgb, int width, int rows, int whitepoint);
    private static native void nSetImageFloat(long ctx, float[] r, float[] g, float[] b, int width, int rows, int whitepoint);
    private static native void nEvalPalettes(long ctx, float[] palettes, int b, int k, int space, int scielab, long[] errFx, long[] counts);
    private static native void nScielabConfigure(long ctx, int dpi, float viewingDistance);
    private static native void nScielabSetFilters(long ctx, float[] filters7, float[] abs3, int taps);
    private static native void nScielabSetImage(long ctx, float[] lab4);
    private static native void nRgbToXyz(long ctx, float[] r, float[] g, float[] b, float[] xyz4);
    private static native void nXyzToScielab(long ctx, float[] xyz4, int width, float[] illuminant, float[] lab4);
    private static native void nQuantize(long ctx, float[] palette, int k, int space, byte[] outRgb, float[] outF32);
    private static native double nDeltaEImages(long ctx, float[] labA, float[] labB, float[] errorImage);
    private static native double nErrorImage(long ctx, byte[] quantizedRgb, float[] errorMap);
    private static native double nErrorImageFloat(long ctx, float[] r, float[] g, float[] b, float[] errorMap);
    private static native int nFindBestQuantization(long ctx, int k, int[] intParams, float[] floatParams, long seed, float[] best,
                                                    double[] bestError, double[] trace, Object listener);
}
