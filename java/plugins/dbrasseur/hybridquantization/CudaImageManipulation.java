package plugins.dbrasseur.hybridquantization;

/**
 * Drop-in CUDA backend for the plugin's hot path: same role as {@code ImageManipulation}
 * (ImageManipulation.java) for findBestQuantization / quantize, bound to libhq_b200.so through
 * the JNI shim java/jni/hq_jni.c.  NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK): delivered
 * for a maintainer with a JDK + Icy; see INTEGRATION.md.
 *
 * The annealing loop, RNG and accept/reject logic stay in Java (SWASA.java is unchanged); only
 * computeQuantizationErrorPopulation (ImageManipulation.java:620-727) and quantize (:770-798)
 * change: one native call per iteration scores the whole population.
 */
public class CudaImageManipulation implements AutoCloseable {
    static { System.loadLibrary("hq_jni"); }

    private long ctx; // hq_ctx*

    public CudaImageManipulation(int device) {
        ctx = nCreate(device); // throws RuntimeException: there is no "pure Java mode" (ImageManipulation.java:79-92)
    }

    /** packed u8 RGB, row-major; replaces the uploads at ImageManipulation.java:451,471-472 */
    public void setImage(byte[] rgb, int width, int rows, int whitepoint) { nSetImage(ctx, rgb, width, rows, whitepoint); }

    /** the converted float image as the plugin holds it: im.getDataXYCAsFloat() (HybridQuantization.java:95-98), channels 0..2 */
    public void setImage(float[][] dataXYC, int width, int rows, int whitepoint) {
        nSetImageFloat(ctx, dataXYC[0], dataXYC[1], dataXYC[2], width, rows, whitepoint);
    }

    /**
     * colors: [population][4*K] as SWASA.generateRandomColors lays them out (SWASA.java:42-50).
     * Returns the costs averageArray(err) + computePenalty(used) of ImageManipulation.java:712.
     */
    public double[] computeQuantizationErrorPopulation(float[][] colors, int nbOfColors, float delta) {
        final int p = colors.length;
        float[] flat = new float[p * 4 * nbOfColors];
        for (int i = 0; i < p; i++) System.arraycopy(colors[i], 0, flat, i * 4 * nbOfColors, 4 * nbOfColors);
        long[] errFx = new long[p];
        long[] counts = new long[p * nbOfColors];
        nEvalPalettes(ctx, flat, p, nbOfColors, 0 /*HQ_SPACE_LAB*/, errFx, counts);
        double[] results = new double[p];
        final long n = nPixels(ctx);
        for (int i = 0; i < p; i++) {
            double penalty = 0;
            for (int k = 0; k < nbOfColors; k++) if (counts[i * nbOfColors + k] == 0) penalty += delta; // SWASA.java:74-82
            results[i] = (errFx[i] * (1.0 / 16777216.0)) / n + penalty;
        }
        return results;
    }

    /** replaces quantize() (ImageManipulation.java:770-798): packed u8 RGB out */
    public byte[] quantize(float[] colors, int nbOfColors) {
        byte[] out = new byte[(int) (3 * nPixels(ctx))];
        nQuantize(ctx, colors, nbOfColors, 0, out);
        return out;
    }

    @Override public void close() { if (ctx != 0) { nDestroy(ctx); ctx = 0; } } // ImageManipulation.close(), :265

    private static native long nCreate(int device);
    private static native void nDestroy(long ctx);
    private static native long nPixels(long ctx);
    private static native void nSetImage(long ctx, byte[] rgb, int width, int rows, int whitepoint);
    private static native void nSetImageFloat(long ctx, float[] r, float[] g, float[] b, int width, int rows, int whitepoint);
    private static native void nEvalPalettes(long ctx, float[] palettes, int b, int k, int space, long[] errFx, long[] counts);
    private static native void nQuantize(long ctx, float[] palette, int k, int space, byte[] outRgb);
}
