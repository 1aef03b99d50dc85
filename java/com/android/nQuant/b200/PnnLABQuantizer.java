package com.android.nQuant.b200;

/**
 * Drop-in for com.android.nQuant.PnnLABQuantizer (reference PnnLABQuantizer.java:17): the CIELAB
 * variant of the same convert(nMaxColors, dither). See PnnQuantizer for the boundary notes.
 */
public class PnnLABQuantizer extends PnnQuantizer {
	public PnnLABQuantizer(int[] argb, int width, int height) {
		super(argb, width, height);
	}

	public PnnLABQuantizer(int[] argb, int width, int height, int device, long rngSeed) {
		super(argb, width, height, device, rngSeed);
	}

	@Override
	protected int kind() {
		return NQuantNative.NQ_KIND_PNNLAB;
	}
}
