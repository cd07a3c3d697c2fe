// Package cubez is a drop-in for github.com/tbogdala/cubez whose per-step pipeline runs on libcubezcuda (B200,
// sm_100a) through cgo.  Exported identifiers, field names and call semantics are the reference's (rigidbody.go,
// colliders.go, contact.go); the math value types are the reference's own pure-Go package.  Added: the batched-world
// handle (World, world.go) and multi-GPU runs (Run, run.go).
//
// There is NO CPU fallback: Init panics without a CUDA device, and every call goes to the library.
//
// Files: cubez.go (context, RigidBody, marshalling) · colliders.go (colliders, CheckForCollisions) · contact.go
// (Contact, ResolveContacts) · world.go (cz_world_*) · run.go (cz_run_*) · real_f64.go / real_f32.go (precision).
//
// Cost model — read before porting a loop: every object-API call (body.Integrate, CheckForCollisions,
// ResolveContacts) is one upload + kernel + download, about 0.7 ms; the reference's body.Integrate is ~100 ns.
// The object API exists for drop-in correctness; throughput comes from the batch entry points (IntegrateBodies,
// CheckCollisionList) and above all from World, which keeps the state on the device.
//
// Compiled by nobody here: the build image has no Go toolchain (tests/test_go_binding.py checks every C.cz_* call
// against include/cubezcuda.h by name and arity instead).
//
//	CGO_CFLAGS=-I<repo>/include CGO_LDFLAGS="-L<repo>/cubez_b200/lib" go build            (float64)
//	... go build -tags cubez_f32    (float32: links libcubezcuda_f32 and needs math.Real = float32, see real_f32.go)
package cubez

/*
#include <stdlib.h>
#include "cubezcuda.h"
*/
import "C"

import (
	"math"
	"runtime"
	"unsafe"

	m "github.com/tbogdala/cubez/math"
)

var ctx *C.cz_ctx

// Init binds the package to a CUDA device.  There is no CPU fallback: it panics without one.
func Init(device int) {
	if int(C.cz_real_size()) != int(unsafe.Sizeof(m.Real(0))) {
		panic("cubez: math.Real and the linked libcubezcuda disagree on the precision (build tag cubez_f32?)")
	}
	if rc := C.cz_init(C.int(device), &ctx); rc != 0 {
		panic("cubez: " + C.GoString(C.cz_last_error(nil)))
	}
}

// Shutdown releases the device context.
func Shutdown() {
	if ctx != nil {
		C.cz_shutdown(ctx)
		ctx = nil
	}
}

// Synchronize waits for everything queued on the context's stream.
func Synchronize() { check(C.cz_ctx_synchronize(ctx)) }

// Stream is the cudaStream_t the context launches on (for interop with other CUDA code).
func Stream() unsafe.Pointer { return C.cz_ctx_stream(ctx) }

func check(rc C.int) {
	if rc != 0 { // the reference has no error returns; failures are panics (contact.go:512-523)
		panic("cubez: " + C.GoString(C.cz_last_error(ctx)))
	}
}

// PinnedReals returns a []m.Real backed by page-locked host memory (cz_host_alloc): buffers of World.StepHost /
// StepRL that the DMA engines read and write directly.  Release with FreePinned.
func PinnedReals(count int) []m.Real {
	var p unsafe.Pointer
	check(C.cz_host_alloc(ctx, C.uint64_t(count)*C.uint64_t(unsafe.Sizeof(C.cz_real(0))), &p))
	return unsafe.Slice((*m.Real)(p), count)
}

// PinnedBytes is PinnedReals for flag arrays (IsAwake, CanSleep).
func PinnedBytes(count int) []uint8 {
	var p unsafe.Pointer
	check(C.cz_host_alloc(ctx, C.uint64_t(count), &p))
	return unsafe.Slice((*uint8)(p), count)
}

// FreePinned releases a PinnedReals buffer.
func FreePinned(buf []m.Real) {
	if len(buf) > 0 {
		check(C.cz_host_free(ctx, unsafe.Pointer(&buf[0])))
	}
}

// FreePinnedBytes releases a PinnedBytes buffer.
func FreePinnedBytes(buf []uint8) {
	if len(buf) > 0 {
		check(C.cz_host_free(ctx, unsafe.Pointer(&buf[0])))
	}
}

// ---------------------------------------------------------------------------------------------
// RigidBody — rigidbody.go:23-101.  Same exported fields; private derived fields mirrored here.
// ---------------------------------------------------------------------------------------------
type RigidBody struct {
	LinearDamping, AngularDamping m.Real
	Position                      m.Vector3
	Orientation                   m.Quat
	Velocity, Acceleration        m.Vector3
	Rotation                      m.Vector3
	InverseInertiaTensor          m.Matrix3
	IsAwake, CanSleep             bool

	inverseInertiaTensorWorld m.Matrix3
	inverseMass, mass         m.Real
	transform                 m.Matrix3x4
	lastFrameAccelleration    m.Vector3
	motion                    m.Real
}

func NewRigidBody() *RigidBody { // rigidbody.go:104-114
	b := new(RigidBody)
	b.Orientation.SetIdentity()
	b.LinearDamping, b.AngularDamping = 0.95, 0.95 // :107-108 assigns defaultLinearDamping to both
	b.Acceleration = m.Vector3{0.0, -9.78, 0.0}
	b.inverseInertiaTensorWorld.SetIdentity()
	b.CanSleep = true
	b.SetAwake(true)
	return b
}
func (b *RigidBody) Clone() *RigidBody                       { c := *b; return &c }
func (b *RigidBody) SetMass(mass m.Real)                     { b.mass, b.inverseMass = mass, 1.0/mass }
func (b *RigidBody) SetInfiniteMass()                        { b.mass, b.inverseMass = 0, 0 }
func (b *RigidBody) HasFiniteMass() bool                     { return b.inverseMass > 0 }
func (b *RigidBody) GetInverseMass() m.Real                  { return b.inverseMass }
func (b *RigidBody) GetTransform() m.Matrix3x4               { return b.transform }
func (b *RigidBody) GetLastFrameAccelleration() m.Vector3    { return b.lastFrameAccelleration }
func (b *RigidBody) GetInverseInertiaTensorWorld() m.Matrix3 { return b.inverseInertiaTensorWorld }
func (b *RigidBody) SetInertiaTensor(t *m.Matrix3)           { b.InverseInertiaTensor = t.Invert() }
func (b *RigidBody) AddVelocity(v *m.Vector3)                { b.Velocity.Add(v) }
func (b *RigidBody) AddRotation(v *m.Vector3)                { b.Rotation.Add(v) }
func (b *RigidBody) ClearAccumulators()                      {} // the accumulators have no writer in the reference (rigidbody.go:86-92)
func (b *RigidBody) GetMass() m.Real {
	if b.inverseMass == 0 {
		return m.MaxValue
	}
	return b.mass
}
func (b *RigidBody) SetAwake(awake bool) { // rigidbody.go:182-192
	if awake {
		b.IsAwake, b.motion = true, 0.6
	} else {
		b.IsAwake = false
		b.Velocity.Clear()
		b.Rotation.Clear()
	}
}

// flat is the cz_bodies SoA of a batch of bodies in C memory (cgo: no Go pointer crosses the boundary).
type flat struct {
	n   int
	c   C.cz_bodies
	mem []unsafe.Pointer
}

func (f *flat) alloc(n, comps int) *C.cz_real {
	p := C.malloc(C.size_t(n*comps) * C.size_t(unsafe.Sizeof(C.cz_real(0))))
	f.mem = append(f.mem, p)
	return (*C.cz_real)(p)
}
func (f *flat) free() {
	for _, p := range f.mem {
		C.free(p)
	}
}
func reals(p *C.cz_real, n int) []m.Real { return unsafe.Slice((*m.Real)(unsafe.Pointer(p)), n) }
func b2u(v bool) uint8 {
	if v {
		return 1
	}
	return 0
}

func gather(bodies []*RigidBody) *flat {
	n := len(bodies)
	f := &flat{n: n}
	f.c.n = C.int32_t(n)
	f.c.position, f.c.orientation, f.c.velocity, f.c.rotation = f.alloc(n, 3), f.alloc(n, 4), f.alloc(n, 3), f.alloc(n, 3)
	f.c.acceleration, f.c.linear_damping, f.c.angular_damping = f.alloc(n, 3), f.alloc(n, 1), f.alloc(n, 1)
	f.c.inverse_inertia_tensor, f.c.inverse_mass, f.c.motion = f.alloc(n, 9), f.alloc(n, 1), f.alloc(n, 1)
	f.c.transform, f.c.inverse_inertia_tensor_world, f.c.last_frame_acceleration = f.alloc(n, 12), f.alloc(n, 9), f.alloc(n, 3)
	aw := C.malloc(C.size_t(n))
	cs := C.malloc(C.size_t(n))
	f.mem = append(f.mem, aw, cs)
	f.c.is_awake, f.c.can_sleep = (*C.uint8_t)(aw), (*C.uint8_t)(cs)
	awake, sleep := unsafe.Slice((*uint8)(aw), n), unsafe.Slice((*uint8)(cs), n)
	for i, b := range bodies {
		copy(reals(f.c.position, 3*n)[3*i:], b.Position[:])
		copy(reals(f.c.orientation, 4*n)[4*i:], b.Orientation[:])
		copy(reals(f.c.velocity, 3*n)[3*i:], b.Velocity[:])
		copy(reals(f.c.rotation, 3*n)[3*i:], b.Rotation[:])
		copy(reals(f.c.acceleration, 3*n)[3*i:], b.Acceleration[:])
		copy(reals(f.c.inverse_inertia_tensor, 9*n)[9*i:], b.InverseInertiaTensor[:])
		copy(reals(f.c.transform, 12*n)[12*i:], b.transform[:])
		copy(reals(f.c.inverse_inertia_tensor_world, 9*n)[9*i:], b.inverseInertiaTensorWorld[:])
		copy(reals(f.c.last_frame_acceleration, 3*n)[3*i:], b.lastFrameAccelleration[:])
		reals(f.c.linear_damping, n)[i], reals(f.c.angular_damping, n)[i] = b.LinearDamping, b.AngularDamping
		reals(f.c.inverse_mass, n)[i], reals(f.c.motion, n)[i] = b.inverseMass, b.motion
		awake[i], sleep[i] = b2u(b.IsAwake), b2u(b.CanSleep)
	}
	return f
}

// scatter mirrors what Integrate / ResolveContacts write back into the Go structs (SURVEY §8b).
func (f *flat) scatter(bodies []*RigidBody) {
	n := f.n
	awake := unsafe.Slice((*uint8)(unsafe.Pointer(f.c.is_awake)), n)
	for i, b := range bodies {
		copy(b.Position[:], reals(f.c.position, 3*n)[3*i:])
		copy(b.Orientation[:], reals(f.c.orientation, 4*n)[4*i:])
		copy(b.Velocity[:], reals(f.c.velocity, 3*n)[3*i:])
		copy(b.Rotation[:], reals(f.c.rotation, 3*n)[3*i:])
		copy(b.transform[:], reals(f.c.transform, 12*n)[12*i:])
		copy(b.inverseInertiaTensorWorld[:], reals(f.c.inverse_inertia_tensor_world, 9*n)[9*i:])
		copy(b.lastFrameAccelleration[:], reals(f.c.last_frame_acceleration, 3*n)[3*i:])
		b.motion, b.IsAwake = reals(f.c.motion, n)[i], awake[i] != 0
	}
}

// Integrate — rigidbody.go:213-259.  The three math.Pow results (:233, :234, :250) are evaluated HERE, in Go, and
// handed to the library, so they are bit-identical to what the reference computes.
func (b *RigidBody) Integrate(duration m.Real) { IntegrateBodies([]*RigidBody{b}, duration) }

// IntegrateBodies is the batch form (one upload, one kernel, one download).
func IntegrateBodies(bodies []*RigidBody, duration m.Real) {
	if len(bodies) == 0 {
		return
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	f := gather(bodies)
	defer f.free()
	n := len(bodies)
	lp, ap := f.alloc(n, 1), f.alloc(n, 1)
	for i, b := range bodies {
		reals(lp, n)[i] = m.Real(math.Pow(float64(b.LinearDamping), float64(duration)))
		reals(ap, n)[i] = m.Real(math.Pow(float64(b.AngularDamping), float64(duration)))
	}
	bias := C.cz_real(m.Real(math.Pow(0.5, float64(duration))))
	check(C.cz_integrate(ctx, &f.c, C.cz_real(duration), lp, ap, &bias))
	f.scatter(bodies)
}

// CalculateDerivedData — rigidbody.go:268-272.
func (b *RigidBody) CalculateDerivedData() {
	f := gather([]*RigidBody{b})
	defer f.free()
	check(C.cz_calculate_derived_data(ctx, &f.c))
	f.scatter([]*RigidBody{b})
}

// MathOp runs one operation of the math layer on the device (cz_math_op; CZ_OP_* codes of cubezcuda.h) — the hook
// the reference's math/*_test.go known answers are checked through.  in holds up to 24 reals, out receives up to 12.
func MathOp(op int, in []m.Real) [12]m.Real {
	var buf [24]m.Real
	var out [12]m.Real
	copy(buf[:], in)
	cin := (*C.cz_real)(C.malloc(C.size_t(24) * C.size_t(unsafe.Sizeof(C.cz_real(0)))))
	cout := (*C.cz_real)(C.malloc(C.size_t(12) * C.size_t(unsafe.Sizeof(C.cz_real(0)))))
	defer C.free(unsafe.Pointer(cin))
	defer C.free(unsafe.Pointer(cout))
	copy(reals(cin, 24), buf[:])
	check(C.cz_math_op(ctx, C.int32_t(op), cin, cout))
	copy(out[:], reals(cout, 12))
	return out
}
