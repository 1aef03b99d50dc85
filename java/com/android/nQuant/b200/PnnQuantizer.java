package com.android.nQuant.b200;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;

/**
 * Drop-in for the quantizer core of com.android.nQuant.PnnQuantizer (reference
 * nQuant.master/src/main/java/com/android/nQuant/PnnQuantizer.java): same constructor role, same
 * convert(nMaxColors, dither) and hasAlpha(), over an ARGB int[] in place of android.graphics.Bitmap.
 * All arithmetic runs in libnquant_b200.so (CUDA, sm_100a); there is no Java fallback.
 *
 * Differences from the reference, all forced by the boundary:
 *  - the constructor takes the pixels the reference obtains with Bitmap.getPixels (PnnQuantizer.java:39-44);
 *  - convert returns the int[] the reference hands to Bitmap.createBitmap (PnnQuantizer.java:455);
 *  - the protected overridables (getQuanFn, pnnquan, nearestColorIndex, closestColorIndex, dither) do
 *    not exist here: a per-pixel Java callback cannot run inside a CUDA kernel.
 *
 * NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no JDK); see INTEGRATION.md.
 */
public class PnnQuantizer implements AutoCloseable {
	protected final int[] pixels;
	protected final int width, height;
	protected final long rngSeed;
	protected boolean hasAlpha;
	protected int[] palette = new int[0];
	private final MemorySegment ctx;

	public PnnQuantizer(int[] argb, int width, int height) {
		this(argb, width, height, 0, 0L);
	}

	/** @param rngSeed seed of the java.util.Random behind PnnLABQuantizer.closestColorIndex
	 *                 (unseeded and static in the reference, PnnLABQuantizer.java:22,467). */
	public PnnQuantizer(int[] argb, int width, int height, int device, long rngSeed) {
		if (argb.length != width * height)
			throw new IllegalArgumentException("argb must hold width*height pixels");
		this.pixels = argb.clone();     // the reference keeps a private copy too (PnnQuantizer.java:42-43)
		this.width = width;
		this.height = height;
		this.rngSeed = rngSeed;
		try {
			this.ctx = (MemorySegment) NQuantNative.nq_create.invokeExact(device);
		} catch (Throwable t) {
			throw new IllegalStateException(t);
		}
		if (ctx.equals(MemorySegment.NULL))
			throw new IllegalStateException("nq_create: " + NQuantNative.lastError());
	}

	protected int kind() {
		return NQuantNative.NQ_KIND_PNN;
	}

	/** Bitmap convert(int nMaxColors, boolean dither) throws Exception (PnnQuantizer.java:409). */
	public int[] convert(int nMaxColors, boolean dither) throws Exception {
		try (Arena arena = Arena.ofConfined()) {
			final long n = (long) width * height;
			MemorySegment in = arena.allocateFrom(ValueLayout.JAVA_INT, pixels);
			MemorySegment out = arena.allocate(ValueLayout.JAVA_INT, n);
			MemorySegment pal = arena.allocate(ValueLayout.JAVA_INT, 256);
			MemorySegment plen = arena.allocate(ValueLayout.JAVA_INT);
			MemorySegment alpha = arena.allocate(ValueLayout.JAVA_INT);
			int rc;
			try {
				rc = (int) NQuantNative.nq_convert.invokeExact(ctx, kind(), in, width, height, nMaxColors, dither ? 1 : 0,
					rngSeed, out, pal, plen, alpha);
			} catch (Throwable t) {
				throw new Exception(t);
			}
			if (rc != 0)
				throw new Exception("nq_convert failed (" + rc + "): " + NQuantNative.lastError());
			hasAlpha = alpha.get(ValueLayout.JAVA_INT, 0) != 0;
			palette = pal.asSlice(0, 4L * plen.get(ValueLayout.JAVA_INT, 0)).toArray(ValueLayout.JAVA_INT);
			return out.toArray(ValueLayout.JAVA_INT);
		}
	}

	/** boolean hasAlpha() (PnnQuantizer.java:458). */
	public boolean hasAlpha() {
		return hasAlpha;
	}

	/** The palette pnnquan produced in the last convert call. */
	public int[] getPalette() {
		return palette.clone();
	}

	@Override
	public void close() {
		try {
			NQuantNative.nq_destroy.invokeExact(ctx);
		} catch (Throwable ignored) {
		}
	}
}
