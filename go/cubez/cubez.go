// Package cubez is a drop-in for github.com/tbogdala/cubez whose per-step pipeline runs on
// libcubezcuda (B200, sm_100a) through cgo.  Exported identifiers, field names and call
// semantics are the reference's (rigidbody.go, colliders.go, contact.go); the math value types
// are the reference's own pure-Go package.  A batched-world handle (World) is added.
//
// SOURCE ONLY: no Go toolchain exists in the build image, so this file has never been compiled.
// It is the binding a maintainer would add; INTEGRATION.md walks through it.
//
// Build: CGO_CFLAGS=-I<repo>/include CGO_LDFLAGS="-L<repo>/cubez_b200/lib -lcubezcuda -lcudart"
//        (float32: -tags cubez_f32, links -lcubezcuda_f32 and needs math.Real = float32).
package cubez

/*
#cgo LDFLAGS: -lcubezcuda -lcudart
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
	if rc := C.cz_init(C.int(device), &ctx); rc != 0 {
		panic("cubez: " + C.GoString(C.cz_last_error(nil)))
	}
}

func check(rc C.int) {
	if rc != 0 { // the reference has no error returns; failures are panics (contact.go:512-523)
		panic("cubez: " + C.GoString(C.cz_last_error(ctx)))
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
	b.LinearDamping, b.AngularDamping = 0.95, 0.95
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
func (b *RigidBody) ClearAccumulators()                      {}
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

// flat is the cz_bodies SoA of a batch of bodies in C memory (no Go pointer crosses the boundary).
type flat struct {
	n  int
	c  C.cz_bodies
	mem []unsafe.Pointer
}

func (f *flat) alloc(n, comps int) *C.cz_real {
	p := C.malloc(C.size_t(n * comps * C.sizeof_cz_real))
	f.mem = append(f.mem, p)
	return (*C.cz_real)(p)
}
func (f *flat) free() {
	for _, p := range f.mem {
		C.free(p)
	}
}
func reals(p *C.cz_real, n int) []m.Real { return unsafe.Slice((*m.Real)(unsafe.Pointer(p)), n) }

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
func b2u(v bool) uint8 {
	if v {
		return 1
	}
	return 0
}

// Integrate — rigidbody.go:213-259.  The three math.Pow results are evaluated HERE, in Go, and
// handed to the library, so they are bit-identical to what the reference computes.
func (b *RigidBody) Integrate(duration m.Real) { IntegrateBodies([]*RigidBody{b}, duration) }

// IntegrateBodies is the batch form (one upload, one kernel, one download).
func IntegrateBodies(bodies []*RigidBody, duration m.Real) {
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
	copy(b.Orientation[:], reals(f.c.orientation, 4)) // Normalize() result
}

// ---------------------------------------------------------------------------------------------
// Colliders — colliders.go:15-71.  CheckAgainst* forward to CheckForCollisions (cz_narrowphase).
// ---------------------------------------------------------------------------------------------
type Collider interface {
	Clone() Collider
	CalculateDerivedData()
	GetBody() *RigidBody
	GetTransform() m.Matrix3x4
	CheckAgainstHalfSpace(plane *CollisionPlane, existingContacts []*Contact) (bool, []*Contact)
	CheckAgainstSphere(sphere *CollisionSphere, existingContacts []*Contact) (bool, []*Contact)
	CheckAgainstCube(secondCube *CollisionCube, existingContacts []*Contact) (bool, []*Contact)
}
type CollisionPlane struct {
	Normal m.Vector3
	Offset m.Real
}
type CollisionCube struct {
	Body      *RigidBody
	Offset    m.Matrix3x4
	transform m.Matrix3x4
	HalfSize  m.Vector3
}
type CollisionSphere struct {
	Body      *RigidBody
	Offset    m.Matrix3x4
	transform m.Matrix3x4
	Radius    m.Real
}

func NewCollisionPlane(n m.Vector3, o m.Real) *CollisionPlane { return &CollisionPlane{n, o} }
func NewCollisionCube(optBody *RigidBody, halfSize m.Vector3) *CollisionCube {
	c := &CollisionCube{Body: optBody, HalfSize: halfSize}
	c.Offset.SetIdentity()
	if c.Body == nil {
		c.Body = NewRigidBody()
	}
	return c
}
func NewCollisionSphere(optBody *RigidBody, radius m.Real) *CollisionSphere {
	s := &CollisionSphere{Body: optBody, Radius: radius}
	s.Offset.SetIdentity()
	if s.Body == nil {
		s.Body = NewRigidBody()
	}
	return s
}
func (p *CollisionPlane) Clone() Collider           { return NewCollisionPlane(p.Normal, p.Offset) }
func (p *CollisionPlane) CalculateDerivedData()     {}
func (p *CollisionPlane) GetBody() *RigidBody       { return nil }
func (p *CollisionPlane) GetTransform() m.Matrix3x4 { var t m.Matrix3x4; t.SetIdentity(); return t }
func (c *CollisionCube) GetBody() *RigidBody        { return c.Body }
func (c *CollisionCube) GetTransform() m.Matrix3x4  { return c.transform }
func (s *CollisionSphere) GetBody() *RigidBody      { return s.Body }
func (s *CollisionSphere) GetTransform() m.Matrix3x4 { return s.transform }
func (c *CollisionCube) Clone() Collider {
	n := NewCollisionCube(c.Body.Clone(), c.HalfSize)
	n.Offset, n.transform = c.Offset, c.transform
	return n
}
func (s *CollisionSphere) Clone() Collider {
	n := NewCollisionSphere(s.Body.Clone(), s.Radius)
	n.Offset, n.transform = s.Offset, s.transform
	return n
}
func derive(body *RigidBody, offset *m.Matrix3x4) (out m.Matrix3x4) { // colliders.go:173-176 / 302-304
	check(C.cz_collider_derive(ctx, 1, (*C.cz_real)(unsafe.Pointer(&body.transform[0])), (*C.cz_real)(unsafe.Pointer(&offset[0])),
		(*C.cz_real)(unsafe.Pointer(&out[0]))))
	return
}
func (c *CollisionCube) CalculateDerivedData()   { c.transform = derive(c.Body, &c.Offset) }
func (s *CollisionSphere) CalculateDerivedData() { s.transform = derive(s.Body, &s.Offset) }

func (p *CollisionPlane) CheckAgainstHalfSpace(_ *CollisionPlane, e []*Contact) (bool, []*Contact) { return false, e }
func (p *CollisionPlane) CheckAgainstSphere(s *CollisionSphere, e []*Contact) (bool, []*Contact)   { return CheckForCollisions(p, s, e) }
func (p *CollisionPlane) CheckAgainstCube(c *CollisionCube, e []*Contact) (bool, []*Contact)       { return CheckForCollisions(p, c, e) }
func (c *CollisionCube) CheckAgainstHalfSpace(p *CollisionPlane, e []*Contact) (bool, []*Contact)  { return CheckForCollisions(c, p, e) }
func (c *CollisionCube) CheckAgainstSphere(s *CollisionSphere, e []*Contact) (bool, []*Contact)    { return CheckForCollisions(c, s, e) }
func (c *CollisionCube) CheckAgainstCube(o *CollisionCube, e []*Contact) (bool, []*Contact)        { return CheckForCollisions(c, o, e) }
func (s *CollisionSphere) CheckAgainstHalfSpace(p *CollisionPlane, e []*Contact) (bool, []*Contact) { return CheckForCollisions(s, p, e) }
func (s *CollisionSphere) CheckAgainstSphere(o *CollisionSphere, e []*Contact) (bool, []*Contact)  { return CheckForCollisions(s, o, e) }
func (s *CollisionSphere) CheckAgainstCube(c *CollisionCube, e []*Contact) (bool, []*Contact)      { return CheckForCollisions(s, c, e) }

// Contact — contact.go:17-51 (public fields).
type Contact struct {
	Bodies                      [2]*RigidBody
	Friction, Restitution       m.Real
	ContactPoint, ContactNormal m.Vector3
	Penetration                 m.Real
}

func NewContact() *Contact { return new(Contact) }

// Check is one ordered (one, two) entry of a pair schedule.
type Check struct{ One, Two Collider }

// CheckForCollisions — colliders.go:720-747.
func CheckForCollisions(one Collider, two Collider, existingContacts []*Contact) (bool, []*Contact) {
	found, contacts := CheckCollisionList([]Check{{one, two}}, existingContacts)
	return found[0], contacts
}

// CheckCollisionList evaluates an ordered list of checks in ONE library call; contacts are
// appended in the order the reference's append calls would produce.
func CheckCollisionList(checks []Check, existing []*Contact) ([]bool, []*Contact) {
	// Flatten the pointer graph to indices (cgo: no Go pointers in C memory): colliders, planes,
	// bodies are numbered in first-use order; see cubez_b200/api.py:check_collision_list for the
	// same marshalling spelled out in Python.  cz_narrowphase fills body indices, point, normal,
	// penetration, friction, restitution; they are mapped back to *RigidBody here.
	panic("marshalling elided in this source-only sketch: see INTEGRATION.md §3")
}

// ResolveContacts — contact.go:208-222.
func ResolveContacts(maxIterations int, contacts []*Contact, duration m.Real) {
	if duration <= 0.0 || len(contacts) == 0 {
		return
	}
	// bodies := unique non-nil bodies of contacts in first-use order -> gather()
	// cz_contacts SoA <- contacts (body0/body1 indices, -1 for nil)
	// C.cz_resolve_contacts(ctx, maxIterations, &cs, &f.c, duration, nil)
	// scatter(): Position, Orientation, Velocity, Rotation, IsAwake, motion (+ transform and world
	// inertia for bodies that were asleep, contact.go:380-382); contacts: Penetration, and Bodies /
	// ContactNormal where Bodies[0] was nil (contact.go:61-65).
	panic("marshalling elided in this source-only sketch: see INTEGRATION.md §3")
}

// ---------------------------------------------------------------------------------------------
// World — the batched-world handle (new API).
// ---------------------------------------------------------------------------------------------
type World struct{ h *C.cz_world }

func NewWorld(nWorlds, bodiesPerWorld, contactsPerWorld int, explicitSchedule bool) *World {
	d := C.cz_world_desc{n_worlds: C.int32_t(nWorlds), bodies_per_world: C.int32_t(bodiesPerWorld), contacts_per_world: C.int32_t(contactsPerWorld)}
	if explicitSchedule {
		d.schedule = C.CZ_SCHED_EXPLICIT
	}
	w := new(World)
	check(C.cz_world_create(ctx, &d, &w.h))
	return w
}

// Step advances every world by n frames of updateCallback (examples/cubedrop.go:69-75).  One cgo
// crossing per call amortises the ~100 ns cgo cost and the kernel launch over n frames.
func (w *World) Step(dt m.Real, n int) C.cz_step_stats {
	var st C.cz_step_stats
	check(C.cz_world_step(w.h, C.cz_real(dt), C.int32_t(n), &st))
	return st
}
func (w *World) ChecksumEnergy() (uint64, float64) {
	var c C.uint64_t
	var e C.double
	check(C.cz_world_checksum_energy(w.h, &c, &e))
	return uint64(c), float64(e)
}
// Observation is the part of the body state StepRL copies back each call; the slices view pinned
// C memory (cz_host_alloc), so neither side copies.
type Observation struct {
	c                                          C.cz_bodies
	Position, Orientation, Velocity, Rotation []m.Real
}

// NewObservation allocates pinned arrays for n bodies (position 3, orientation 4, velocity 3, rotation 3).
func NewObservation(n int) *Observation {
	o := new(Observation)
	o.c.n = C.int32_t(n)
	pin := func(count int) *C.cz_real {
		var p unsafe.Pointer
		check(C.cz_host_alloc(ctx, C.uint64_t(count)*C.uint64_t(unsafe.Sizeof(C.cz_real(0))), &p))
		return (*C.cz_real)(p)
	}
	o.c.position, o.c.orientation, o.c.velocity, o.c.rotation = pin(3*n), pin(4*n), pin(3*n), pin(3*n)
	o.Position, o.Orientation = reals(o.c.position, 3*n), reals(o.c.orientation, 4*n)
	o.Velocity, o.Rotation = reals(o.c.velocity, 3*n), reals(o.c.rotation, 3*n)
	return o
}

// PinnedReals returns a pinned []m.Real (action buffers of StepRL).
func PinnedReals(count int) []m.Real {
	var p unsafe.Pointer
	check(C.cz_host_alloc(ctx, C.uint64_t(count)*C.uint64_t(unsafe.Sizeof(C.cz_real(0))), &p))
	return reals((*C.cz_real)(p), count)
}

// StepRL is the RL loop's frame: AddVelocity(addVelocity[3i:]) and AddRotation(addRotation[3i:]) on every
// body (rigidbody.go:195-202; either slice may be nil), n frames on the device, then obs filled.
func (w *World) StepRL(addVelocity, addRotation []m.Real, obs *Observation, dt m.Real, n int) C.cz_step_stats {
	var st C.cz_step_stats
	var av, ar *C.cz_real
	if addVelocity != nil {
		av = (*C.cz_real)(unsafe.Pointer(&addVelocity[0]))
	}
	if addRotation != nil {
		ar = (*C.cz_real)(unsafe.Pointer(&addRotation[0]))
	}
	var o *C.cz_bodies
	if obs != nil {
		o = &obs.c
	}
	check(C.cz_world_step_rl(w.h, av, ar, o, C.cz_real(dt), C.int32_t(n), &st))
	return st
}

// SetMaterials replaces the hard-wired `c.Friction = 0.9` / `c.Restitution = 0.1` test constants
// (colliders.go:199-202 and five more FIXME sites) by a table lookup: friction and restitution are
// nMaterials x nMaterials tables, row = material of CheckForCollisions' `one`, column = `two`;
// bodyMaterial holds one id per body of every world, planeMaterial one id per plane.
// nMaterials = 0 restores the constants.
func (w *World) SetMaterials(nMaterials int, friction, restitution []m.Real, nWorlds int, bodyMaterial, planeMaterial []int32) {
	var f, r *C.cz_real
	var bm, pm *C.int32_t
	if nMaterials > 0 {
		f, r = (*C.cz_real)(unsafe.Pointer(&friction[0])), (*C.cz_real)(unsafe.Pointer(&restitution[0]))
	}
	if bodyMaterial != nil {
		bm = (*C.int32_t)(unsafe.Pointer(&bodyMaterial[0]))
	}
	if planeMaterial != nil {
		pm = (*C.int32_t)(unsafe.Pointer(&planeMaterial[0]))
	}
	check(C.cz_world_set_materials(w.h, C.int32_t(nMaterials), f, r, 0, C.int32_t(nWorlds), bm, pm))
}

// ExportGL fills float32 Location (3 per body) and LocalRotation (4 per body: W, V[0], V[1], V[2]) of every body —
// the per-frame SetGlVector3 / SetGlQuat copy of examples/cubedrop.go:35-37 and examples/exampleapp.go:146-159,
// converted on the device.  model may be nil; otherwise it receives the body transform as a column-major 4x4.
func (w *World) ExportGL(nWorlds int, location, rotation, model []float32) {
	var l, q, md *C.float
	if location != nil {
		l = (*C.float)(unsafe.Pointer(&location[0]))
	}
	if rotation != nil {
		q = (*C.float)(unsafe.Pointer(&rotation[0]))
	}
	if model != nil {
		md = (*C.float)(unsafe.Pointer(&model[0]))
	}
	check(C.cz_world_export_gl(w.h, 0, C.int32_t(nWorlds), l, q, md, 0))
}

func (w *World) Close() { C.cz_world_destroy(w.h) }
