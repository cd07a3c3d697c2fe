//go:build !cubez_f32

package cubez

// float64 build (the reference's default `type Real float64`, math/math.go:23): links libcubezcuda.so.

/*
#cgo LDFLAGS: -lcubezcuda -lcudart
*/
import "C"

// RealIsFloat32 reports which precision this build of the package was compiled for.
const RealIsFloat32 = false
