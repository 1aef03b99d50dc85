package com.android.nQuant.b200;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.foreign.ValueLayout;
import java.lang.invoke.MethodHandle;

/**
 * java.lang.foreign (JDK 22+) binding of include/nquant_b200.h -- the C ABI of libnquant_b200.so.
 * No JNI glue is needed: every entry point takes plain pointers and sizes.
 *
 * NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (it has no JDK). The same ABI is exercised through
 * ctypes by nquant_android_b200/_lib.py and tests/test_abi.py; the descriptors below mirror that
 * file one to one.
 */
final class NQuantNative {
	static final int NQ_KIND_PNN = 0;      // com.android.nQuant.PnnQuantizer
	static final int NQ_KIND_PNNLAB = 1;   // com.android.nQuant.PnnLABQuantizer

	private static final Linker LINKER = Linker.nativeLinker();
	private static final SymbolLookup LIB = SymbolLookup.libraryLookup(
		System.getProperty("nquant.b200.lib", "libnquant_b200.so"), Arena.global());

	private static MethodHandle h(String name, FunctionDescriptor fd) {
		return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
	}

	private static final ValueLayout.OfInt I = ValueLayout.JAVA_INT;
	private static final ValueLayout.OfLong L = ValueLayout.JAVA_LONG;
	private static final java.lang.foreign.AddressLayout P = ValueLayout.ADDRESS;

	// nq_ctx* nq_create(int device); void nq_destroy(nq_ctx*); const char* nq_last_error(void)
	static final MethodHandle nq_create = h("nq_create", FunctionDescriptor.of(P, I));
	static final MethodHandle nq_destroy = h("nq_destroy", FunctionDescriptor.ofVoid(P));
	static final MethodHandle nq_last_error = h("nq_last_error", FunctionDescriptor.of(P));
	// int nq_convert(ctx, kind, argb_in, width, height, n_max_colors, dither, rng_seed, argb_out, palette_out, palette_len, has_alpha)
	static final MethodHandle nq_convert = h("nq_convert", FunctionDescriptor.of(I, P, I, P, I, I, I, I, L, P, P, P, P));
	// int nq_convert_batch(ctx, kind, argb_in, n_images, width, height, n_max_colors, dither, rng_seeds, argb_out, palettes_out, palette_lens, has_alpha)
	static final MethodHandle nq_convert_batch = h("nq_convert_batch", FunctionDescriptor.of(I, P, I, P, I, I, I, I, I, P, P, P, P, P));
	// int nq_set_spec_dither(ctx, on, segment, warmup): opt-in speculative segment-parallel error diffusion (DESIGN.md 7.1)
	static final MethodHandle nq_set_spec_dither = h("nq_set_spec_dither", FunctionDescriptor.of(I, P, I, I, I));

	static String lastError() {
		try {
			MemorySegment s = (MemorySegment) nq_last_error.invokeExact();
			return s.reinterpret(4096).getString(0);
		} catch (Throwable t) {
			return t.toString();
		}
	}

	private NQuantNative() {
	}
}
