//go:build cubez_f32

package cubez

// float32 build, selected with `-tags cubez_f32`: links libcubezcuda_f32.so (compiled with -DCUBEZ_REAL_FLOAT) and
// defines CUBEZ_REAL_FLOAT for cgo so that cz_real is float in cubezcuda.h.  The reference's math package must be
// switched too — its only precision switch is editing `type Real float64` (math/math.go:23), and
// `MaxValue = Real(math.MaxFloat64)` (math/math.go:32) must become math.MaxFloat32 or it does not compile
// (SURVEY Appendix D).  Init() panics when the two sides disagree.

/*
#cgo CFLAGS: -DCUBEZ_REAL_FLOAT
#cgo LDFLAGS: -lcubezcuda_f32 -lcudart
*/
import "C"

// RealIsFloat32 reports which precision this build of the package was compiled for.
const RealIsFloat32 = true
