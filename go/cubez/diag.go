package cubez

/*
#include <stdlib.h>
#include "cubezcuda.h"
*/
import "C"

import (
	"unsafe"

	m "github.com/tbogdala/cubez/math"
)

// Microbenchmarks and test hooks of the library (cubezcuda.h "microbench / diagnostics").

// BenchIntegrate — cz_bench_integrate: Integrate + CalculateDerivedData over n device-resident free bodies (BASELINE
// config 5), CUDA-event timed; returns the average ms per step and a checksum of the first 4 096 bodies.
func BenchIntegrate(n int64, seed uint64, warmup, steps int, dt m.Real) (avgMs float32, checksum uint64) {
	var ms C.float
	var cks C.uint64_t
	check(C.cz_bench_integrate(ctx, C.int64_t(n), C.uint64_t(seed), C.int32_t(warmup), C.int32_t(steps), C.cz_real(dt), &ms, &cks))
	return float32(ms), uint64(cks)
}

// BenchFP64Rate — cz_bench_fp64_rate: measured thread-level FP64 instructions per second of the device.
func BenchFP64Rate() float64 {
	var r C.double
	check(C.cz_bench_fp64_rate(ctx, &r))
	return float64(r)
}

// BenchBroadphase — cz_bench_broadphase: the sort-based broadphase on n unit spheres at `fill` volume fraction.
func BenchBroadphase(n int64, seed uint64, fill float64, warmup, steps int) (avgMs float32, pairs int64, sortMs float32) {
	var ms, sms C.float
	var np C.int64_t
	check(C.cz_bench_broadphase(ctx, C.int64_t(n), C.uint64_t(seed), C.double(fill), C.int32_t(warmup), C.int32_t(steps), &ms, &np, &sms))
	return float32(ms), int64(np), float32(sms)
}

// BroadphasePairs — cz_broadphase_pairs: candidate pairs (i, j) of n bounding spheres (centers: 3 per sphere).
func BroadphasePairs(centers, radii []m.Real, capacity int64) [][2]int32 {
	n := len(radii)
	rs := unsafe.Sizeof(C.cz_real(0))
	cp, rp, pp := cbuf(3*n, rs), cbuf(n, rs), cbuf(int(2*capacity), 4)
	defer C.free(cp)
	defer C.free(rp)
	defer C.free(pp)
	copy(reals((*C.cz_real)(cp), 3*n), centers)
	copy(reals((*C.cz_real)(rp), n), radii)
	var cnt C.int64_t
	check(C.cz_broadphase_pairs(ctx, C.int64_t(n), (*C.cz_real)(cp), (*C.cz_real)(rp), C.int64_t(capacity), (*C.int32_t)(pp), &cnt))
	flat := unsafe.Slice((*int32)(pp), 2*int(cnt))
	out := make([][2]int32, int(cnt))
	for k := range out {
		out[k] = [2]int32{flat[2*k], flat[2*k+1]}
	}
	return out
}

// SortPairsU32 / SortPairsU64 — the hand-written LSD radix sort on host buffers (in place, stable).
func SortPairsU32(keys, vals []uint32) {
	n := len(keys)
	kp, vp := cbuf(n, 4), cbuf(n, 4)
	defer C.free(kp)
	defer C.free(vp)
	copy(unsafe.Slice((*uint32)(kp), n), keys)
	copy(unsafe.Slice((*uint32)(vp), n), vals)
	check(C.cz_sort_pairs_u32(ctx, C.int64_t(n), (*C.uint32_t)(kp), (*C.uint32_t)(vp)))
	copy(keys, unsafe.Slice((*uint32)(kp), n))
	copy(vals, unsafe.Slice((*uint32)(vp), n))
}
func SortPairsU64(keys []uint64, vals []uint32, bits int) {
	n := len(keys)
	kp, vp := cbuf(n, 8), cbuf(n, 4)
	defer C.free(kp)
	defer C.free(vp)
	copy(unsafe.Slice((*uint64)(kp), n), keys)
	copy(unsafe.Slice((*uint32)(vp), n), vals)
	check(C.cz_sort_pairs_u64(ctx, C.int64_t(n), (*C.uint64_t)(kp), (*C.uint32_t)(vp), C.int32_t(bits)))
	copy(keys, unsafe.Slice((*uint64)(kp), n))
	copy(vals, unsafe.Slice((*uint32)(vp), n))
}
