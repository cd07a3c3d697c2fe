package cubez

/*
#include <stdlib.h>
#include "cubezcuda.h"
*/
import "C"

import (
	"math"
	"unsafe"

	m "github.com/tbogdala/cubez/math"
)

// ---------------------------------------------------------------------------------------------
// World — the batched-world handle (new API, cz_world_*): nWorlds independent worlds resident on one GPU, each
// stepped by the frame loop of examples/cubedrop.go:69-75 (or an explicit ordered check list, examples/ballistic.go).
// Body i of a world is collider i of that world.
// ---------------------------------------------------------------------------------------------

// Pair schedules and world flags (cubezcuda.h).
const (
	SchedAllPairsOrdered = int(C.CZ_SCHED_ALL_PAIRS_ORDERED) // examples/cubedrop.go:47-64
	SchedExplicit        = int(C.CZ_SCHED_EXPLICIT)          // examples/ballistic.go:47-97
	WorldBroadphase      = int(C.CZ_WORLD_BROADPHASE)        // one large world through the sort-based broadphase
	WorldFused           = int(C.CZ_WORLD_FUSED)
	WorldNoFused         = int(C.CZ_WORLD_NO_FUSED)
)

// StepStats is cz_step_stats.
type StepStats struct {
	Steps, Contacts, PosIterations, VelIterations, Checks, KernelLaunches int64
	MaxContacts                                                            int
	DeviceMs                                                               float32
}

func statsOf(st *C.cz_step_stats) StepStats {
	return StepStats{int64(st.steps), int64(st.contacts), int64(st.pos_iterations), int64(st.vel_iterations), int64(st.checks),
		int64(st.kernel_launches), int(st.max_contacts), float32(st.device_ms)}
}

// HostBodies is a cz_bodies SoA whose arrays live in page-locked C memory (cz_host_alloc) and are viewed as Go slices:
// the buffer type of uploads, downloads, StepHost (state in / state out) and StepRL (observations).  Only the fields
// named at construction exist; the others stay nil (= "skip that field" for the library).
type HostBodies struct {
	c                                                                        C.cz_bodies
	N                                                                        int
	Position, Orientation, Velocity, Rotation, Acceleration                  []m.Real
	LinearDamping, AngularDamping, InverseInertiaTensor, InverseMass, Motion []m.Real
	Transform, InverseInertiaTensorWorld, LastFrameAcceleration              []m.Real
	IsAwake, CanSleep                                                        []uint8
}

// AllBodyFields names every array of HostBodies; ObservationFields is what an RL loop usually reads back.
var (
	AllBodyFields = []string{"position", "orientation", "velocity", "rotation", "acceleration", "linear_damping", "angular_damping",
		"inverse_inertia_tensor", "inverse_mass", "motion", "transform", "inverse_inertia_tensor_world", "last_frame_acceleration", "is_awake", "can_sleep"}
	ObservationFields = []string{"position", "orientation", "velocity", "rotation"}
)

// NewHostBodies allocates pinned arrays for n bodies (all fields when none is named).
func NewHostBodies(n int, fields ...string) *HostBodies {
	if len(fields) == 0 {
		fields = AllBodyFields
	}
	h := &HostBodies{N: n}
	h.c.n = C.int32_t(n)
	pin := func(count int) (*C.cz_real, []m.Real) {
		s := PinnedReals(count)
		return (*C.cz_real)(unsafe.Pointer(&s[0])), s
	}
	for _, f := range fields {
		switch f {
		case "position":
			h.c.position, h.Position = pin(3 * n)
		case "orientation":
			h.c.orientation, h.Orientation = pin(4 * n)
		case "velocity":
			h.c.velocity, h.Velocity = pin(3 * n)
		case "rotation":
			h.c.rotation, h.Rotation = pin(3 * n)
		case "acceleration":
			h.c.acceleration, h.Acceleration = pin(3 * n)
		case "linear_damping":
			h.c.linear_damping, h.LinearDamping = pin(n)
		case "angular_damping":
			h.c.angular_damping, h.AngularDamping = pin(n)
		case "inverse_inertia_tensor":
			h.c.inverse_inertia_tensor, h.InverseInertiaTensor = pin(9 * n)
		case "inverse_mass":
			h.c.inverse_mass, h.InverseMass = pin(n)
		case "motion":
			h.c.motion, h.Motion = pin(n)
		case "transform":
			h.c.transform, h.Transform = pin(12 * n)
		case "inverse_inertia_tensor_world":
			h.c.inverse_inertia_tensor_world, h.InverseInertiaTensorWorld = pin(9 * n)
		case "last_frame_acceleration":
			h.c.last_frame_acceleration, h.LastFrameAcceleration = pin(3 * n)
		case "is_awake":
			h.IsAwake = PinnedBytes(n)
			h.c.is_awake = (*C.uint8_t)(unsafe.Pointer(&h.IsAwake[0]))
		case "can_sleep":
			h.CanSleep = PinnedBytes(n)
			h.c.can_sleep = (*C.uint8_t)(unsafe.Pointer(&h.CanSleep[0]))
		default:
			panic("cubez: unknown body field " + f)
		}
	}
	return h
}

// Free releases the pinned arrays.
func (h *HostBodies) Free() {
	for _, s := range [][]m.Real{h.Position, h.Orientation, h.Velocity, h.Rotation, h.Acceleration, h.LinearDamping, h.AngularDamping,
		h.InverseInertiaTensor, h.InverseMass, h.Motion, h.Transform, h.InverseInertiaTensorWorld, h.LastFrameAcceleration} {
		FreePinned(s)
	}
	FreePinnedBytes(h.IsAwake)
	FreePinnedBytes(h.CanSleep)
	*h = HostBodies{}
}

// Set copies the primary state of a RigidBody into slot i (fields that exist).
func (h *HostBodies) Set(i int, b *RigidBody) {
	put := func(dst []m.Real, comps int, src []m.Real) {
		if dst != nil {
			copy(dst[comps*i:], src)
		}
	}
	put(h.Position, 3, b.Position[:])
	put(h.Orientation, 4, b.Orientation[:])
	put(h.Velocity, 3, b.Velocity[:])
	put(h.Rotation, 3, b.Rotation[:])
	put(h.Acceleration, 3, b.Acceleration[:])
	put(h.InverseInertiaTensor, 9, b.InverseInertiaTensor[:])
	put(h.Transform, 12, b.transform[:])
	put(h.InverseInertiaTensorWorld, 9, b.inverseInertiaTensorWorld[:])
	put(h.LastFrameAcceleration, 3, b.lastFrameAccelleration[:])
	put(h.LinearDamping, 1, []m.Real{b.LinearDamping})
	put(h.AngularDamping, 1, []m.Real{b.AngularDamping})
	put(h.InverseMass, 1, []m.Real{b.inverseMass})
	put(h.Motion, 1, []m.Real{b.motion})
	if h.IsAwake != nil {
		h.IsAwake[i] = b2u(b.IsAwake)
	}
	if h.CanSleep != nil {
		h.CanSleep[i] = b2u(b.CanSleep)
	}
}

// Get copies slot i back into a RigidBody (what a frame writes: SURVEY §8b).
func (h *HostBodies) Get(i int, b *RigidBody) {
	get := func(dst []m.Real, comps int, src []m.Real) {
		if src != nil {
			copy(dst, src[comps*i:comps*i+comps])
		}
	}
	get(b.Position[:], 3, h.Position)
	get(b.Orientation[:], 4, h.Orientation)
	get(b.Velocity[:], 3, h.Velocity)
	get(b.Rotation[:], 3, h.Rotation)
	get(b.transform[:], 12, h.Transform)
	get(b.inverseInertiaTensorWorld[:], 9, h.InverseInertiaTensorWorld)
	get(b.lastFrameAccelleration[:], 3, h.LastFrameAcceleration)
	if h.Motion != nil {
		b.motion = h.Motion[i]
	}
	if h.IsAwake != nil {
		b.IsAwake = h.IsAwake[i] != 0
	}
}

// HostColliders is a cz_colliders SoA in C memory.
type HostColliders struct {
	c                                   C.cz_colliders
	N                                   int
	Shape, Body                         []int32
	Offset, Transform, HalfSize, Radius []m.Real
	mem                                 []unsafe.Pointer
}

// NewHostColliders allocates arrays for n colliders.
func NewHostColliders(n int) *HostColliders {
	h := &HostColliders{N: n}
	rs := unsafe.Sizeof(C.cz_real(0))
	sp, bp := cbuf(n, 4), cbuf(n, 4)
	op, tp, hp, rp := cbuf(12*n, rs), cbuf(12*n, rs), cbuf(3*n, rs), cbuf(n, rs)
	h.mem = []unsafe.Pointer{sp, bp, op, tp, hp, rp}
	h.c.n = C.int32_t(n)
	h.c.shape, h.c.body = (*C.int32_t)(sp), (*C.int32_t)(bp)
	h.c.offset, h.c.transform, h.c.half_size, h.c.radius = (*C.cz_real)(op), (*C.cz_real)(tp), (*C.cz_real)(hp), (*C.cz_real)(rp)
	h.Shape, h.Body = unsafe.Slice((*int32)(sp), n), unsafe.Slice((*int32)(bp), n)
	h.Offset, h.Transform = reals(h.c.offset, 12*n), reals(h.c.transform, 12*n)
	h.HalfSize, h.Radius = reals(h.c.half_size, 3*n), reals(h.c.radius, n)
	return h
}

// Free releases the arrays.
func (h *HostColliders) Free() {
	for _, p := range h.mem {
		C.free(p)
	}
	*h = HostColliders{}
}

// Set fills slot i from a collider object (nil: an unused slot, CZ_SHAPE_NONE).
func (h *HostColliders) Set(i int, c Collider) {
	h.Body[i] = int32(i)
	var ident m.Matrix3x4
	ident.SetIdentity()
	copy(h.Offset[12*i:], ident[:])
	copy(h.Transform[12*i:], ident[:])
	h.HalfSize[3*i], h.HalfSize[3*i+1], h.HalfSize[3*i+2], h.Radius[i] = 0, 0, 0, 0
	switch v := c.(type) {
	case *CollisionCube:
		h.Shape[i] = C.CZ_SHAPE_CUBE
		copy(h.Offset[12*i:], v.Offset[:])
		copy(h.Transform[12*i:], v.transform[:])
		copy(h.HalfSize[3*i:], v.HalfSize[:])
	case *CollisionSphere:
		h.Shape[i] = C.CZ_SHAPE_SPHERE
		copy(h.Offset[12*i:], v.Offset[:])
		copy(h.Transform[12*i:], v.transform[:])
		h.Radius[i] = v.Radius
	default:
		h.Shape[i] = C.CZ_SHAPE_NONE
	}
}

// World is a handle on nWorlds device-resident worlds of bodiesPerWorld body slots each.
type World struct {
	h                                         *C.cz_world
	NWorlds, BodiesPerWorld, ContactsPerWorld int
}

// NewWorld — cz_world_create.  schedule: SchedAllPairsOrdered | SchedExplicit; flags: World* constants.
func NewWorld(nWorlds, bodiesPerWorld, contactsPerWorld, schedule, flags int) *World {
	d := C.cz_world_desc{n_worlds: C.int32_t(nWorlds), bodies_per_world: C.int32_t(bodiesPerWorld),
		contacts_per_world: C.int32_t(contactsPerWorld), schedule: C.int32_t(schedule), flags: C.int32_t(flags)}
	w := &World{NWorlds: nWorlds, BodiesPerWorld: bodiesPerWorld, ContactsPerWorld: contactsPerWorld}
	check(C.cz_world_create(ctx, &d, &w.h))
	return w
}

// Close — cz_world_destroy.
func (w *World) Close() {
	if w.h != nil {
		C.cz_world_destroy(w.h)
		w.h = nil
	}
}

// UploadBodies — cz_world_upload_bodies for worlds [firstWorld, firstWorld + b.N/BodiesPerWorld).  derive: run
// CalculateDerivedData on the device instead of uploading transform / world inertia.
func (w *World) UploadBodies(firstWorld int, b *HostBodies, derive bool) {
	check(C.cz_world_upload_bodies(w.h, C.int32_t(firstWorld), C.int32_t(b.N/w.BodiesPerWorld), &b.c, C.int32_t(b2u(derive))))
}

// UploadColliders — cz_world_upload_colliders.
func (w *World) UploadColliders(firstWorld int, c *HostColliders, derive bool) {
	check(C.cz_world_upload_colliders(w.h, C.int32_t(firstWorld), C.int32_t(c.N/w.BodiesPerWorld), &c.c, C.int32_t(b2u(derive))))
}

// UploadObjects uploads worlds built from the reference's own objects: colliders holds NWorlds*BodiesPerWorld entries,
// world-major (nil = unused slot); body i is colliders[i].GetBody().  Derived data is recomputed on the device.  The
// three math.Pow factors of every body (rigidbody.go:233, :234, :250) are evaluated here in Go for duration dt and
// handed to the library (SetPow), so a Go host integrates bit-identically to the reference.
func (w *World) UploadObjects(colliders []Collider, dt m.Real) {
	n := len(colliders)
	hb, hc := NewHostBodies(n), NewHostColliders(n)
	defer hb.Free()
	defer hc.Free()
	lin, ang := make([]m.Real, n), make([]m.Real, n)
	blank := NewRigidBody()
	for i, c := range colliders {
		b := blank
		if c != nil && c.GetBody() != nil {
			b = c.GetBody()
		}
		hb.Set(i, b)
		hc.Set(i, c)
		lin[i] = m.Real(math.Pow(float64(b.LinearDamping), float64(dt)))
		ang[i] = m.Real(math.Pow(float64(b.AngularDamping), float64(dt)))
	}
	w.UploadBodies(0, hb, true)
	w.UploadColliders(0, hc, true)
	w.SetPow(dt, lin, ang, m.Real(math.Pow(0.5, float64(dt))))
}

// UploadPlanes — cz_world_upload_planes (shared by all worlds; at most 8).
func (w *World) UploadPlanes(planes []*CollisionPlane) {
	n := len(planes)
	rs := unsafe.Sizeof(C.cz_real(0))
	np, op := cbuf(3*n, rs), cbuf(n, rs)
	defer C.free(np)
	defer C.free(op)
	p := C.cz_planes{n: C.int32_t(n), normal: (*C.cz_real)(np), offset: (*C.cz_real)(op)}
	for i, pl := range planes {
		copy(reals(p.normal, 3*n)[3*i:], pl.Normal[:])
		reals(p.offset, n)[i] = pl.Offset
	}
	check(C.cz_world_upload_planes(w.h, &p))
}

// UploadSchedule — cz_world_upload_schedule: the explicit ordered check list of a SchedExplicit world
// (one[k], two[k] >= 0: body index; < 0: plane -(p+1)); upload the planes first.
func (w *World) UploadSchedule(one, two []int32) {
	n := len(one)
	a, b := cbuf(n, 4), cbuf(n, 4)
	defer C.free(a)
	defer C.free(b)
	copy(unsafe.Slice((*int32)(a), n), one)
	copy(unsafe.Slice((*int32)(b), n), two)
	check(C.cz_world_upload_schedule(w.h, C.int32_t(n), (*C.int32_t)(a), (*C.int32_t)(b)))
}

// SetActivation — cz_world_set_activation: activeFrom[i] = first frame body i takes part in (nil: 0);
// integrate[i] == 0: body i is never integrated (the ballistic backboard; nil: all integrated).
func (w *World) SetActivation(firstWorld, nWorlds int, activeFrom []int32, integrate []uint8) {
	var a *C.int32_t
	var g *C.uint8_t
	if activeFrom != nil {
		p := cbuf(len(activeFrom), 4)
		defer C.free(p)
		copy(unsafe.Slice((*int32)(p), len(activeFrom)), activeFrom)
		a = (*C.int32_t)(p)
	}
	if integrate != nil {
		p := cbuf(len(integrate), 1)
		defer C.free(p)
		copy(unsafe.Slice((*uint8)(p), len(integrate)), integrate)
		g = (*C.uint8_t)(p)
	}
	check(C.cz_world_set_activation(w.h, C.int32_t(firstWorld), C.int32_t(nWorlds), a, g))
}

// SetPow — cz_world_set_pow: the host language's own math.Pow results for duration dt, one pair per body of the
// handle, plus bias = Pow(0.5, dt).
func (w *World) SetPow(dt m.Real, linPow, angPow []m.Real, bias m.Real) {
	n := len(linPow)
	rs := unsafe.Sizeof(C.cz_real(0))
	lp, ap := cbuf(n, rs), cbuf(n, rs)
	defer C.free(lp)
	defer C.free(ap)
	copy(reals((*C.cz_real)(lp), n), linPow)
	copy(reals((*C.cz_real)(ap), n), angPow)
	check(C.cz_world_set_pow(w.h, C.cz_real(dt), (*C.cz_real)(lp), (*C.cz_real)(ap), C.cz_real(bias)))
}

// AddForces — cz_world_add_forces: forceAccum += force[3i:], torqueAccum += torque[3i:] for every body of worlds
// [firstWorld, firstWorld+nWorlds) (either slice may be nil) — the writer the reference's accumulators never had
// (rigidbody.go:86-92, read at :219-223, cleared at :206).
func (w *World) AddForces(firstWorld, nWorlds int, force, torque []m.Real) {
	rs := unsafe.Sizeof(C.cz_real(0))
	var f, t *C.cz_real
	if force != nil {
		p := cbuf(len(force), rs)
		defer C.free(p)
		copy(reals((*C.cz_real)(p), len(force)), force)
		f = (*C.cz_real)(p)
	}
	if torque != nil {
		p := cbuf(len(torque), rs)
		defer C.free(p)
		copy(reals((*C.cz_real)(p), len(torque)), torque)
		t = (*C.cz_real)(p)
	}
	check(C.cz_world_add_forces(w.h, C.int32_t(firstWorld), C.int32_t(nWorlds), f, t))
}

// SetStepIndex — cz_world_set_step_index (activation and episodes count frames from it).
func (w *World) SetStepIndex(step int64) { check(C.cz_world_set_step_index(w.h, C.int64_t(step))) }

// SetEpisodes — cz_world_set_episodes: snapshot now; world k is at frame phase0[k] of an episode of `length` frames
// and is restored to the snapshot whenever its phase wraps.  length <= 0 disables.
func (w *World) SetEpisodes(length int, phase0 []int32) {
	var p *C.int32_t
	if phase0 != nil {
		q := cbuf(len(phase0), 4)
		defer C.free(q)
		copy(unsafe.Slice((*int32)(q), len(phase0)), phase0)
		p = (*C.int32_t)(q)
	}
	check(C.cz_world_set_episodes(w.h, C.int32_t(length), p))
}

// SetMaterials — cz_world_set_materials: replaces the hard-wired `c.Friction = 0.9` / `c.Restitution = 0.1` test
// constants (colliders.go:199-202 and five more FIXME sites) by a table lookup: friction and restitution are
// nMaterials x nMaterials tables, row = material of CheckForCollisions' `one`, column = `two`; bodyMaterial holds one id
// per body of worlds [firstWorld, firstWorld+nWorlds), planeMaterial one id per plane.  nMaterials = 0 restores the constants.
func (w *World) SetMaterials(nMaterials int, friction, restitution []m.Real, firstWorld, nWorlds int, bodyMaterial, planeMaterial []int32) {
	rs := unsafe.Sizeof(C.cz_real(0))
	var f, r *C.cz_real
	var bm, pm *C.int32_t
	if nMaterials > 0 {
		fp, rp := cbuf(len(friction), rs), cbuf(len(restitution), rs)
		defer C.free(fp)
		defer C.free(rp)
		copy(reals((*C.cz_real)(fp), len(friction)), friction)
		copy(reals((*C.cz_real)(rp), len(restitution)), restitution)
		f, r = (*C.cz_real)(fp), (*C.cz_real)(rp)
	}
	if bodyMaterial != nil {
		p := cbuf(len(bodyMaterial), 4)
		defer C.free(p)
		copy(unsafe.Slice((*int32)(p), len(bodyMaterial)), bodyMaterial)
		bm = (*C.int32_t)(p)
	}
	if planeMaterial != nil {
		p := cbuf(len(planeMaterial), 4)
		defer C.free(p)
		copy(unsafe.Slice((*int32)(p), len(planeMaterial)), planeMaterial)
		pm = (*C.int32_t)(p)
	}
	check(C.cz_world_set_materials(w.h, C.int32_t(nMaterials), f, r, C.int32_t(firstWorld), C.int32_t(nWorlds), bm, pm))
}

// Step advances every world by n frames of updateCallback (examples/cubedrop.go:69-75) and waits for the counters.  One
// cgo crossing per call amortises the cgo cost and the kernel launches over n frames.
func (w *World) Step(dt m.Real, n int) StepStats {
	var st C.cz_step_stats
	check(C.cz_world_step(w.h, C.cz_real(dt), C.int32_t(n), &st))
	return statsOf(&st)
}

// StepAsync enqueues n frames and returns at once; device-side errors surface at the next Synchronize / download.
func (w *World) StepAsync(dt m.Real, n int) { check(C.cz_world_step(w.h, C.cz_real(dt), C.int32_t(n), nil)) }

// Synchronize — cz_world_synchronize.
func (w *World) Synchronize() { check(C.cz_world_synchronize(w.h)) }

// DownloadBodies — cz_world_download_bodies into the fields `out` holds.
func (w *World) DownloadBodies(firstWorld int, out *HostBodies) {
	check(C.cz_world_download_bodies(w.h, C.int32_t(firstWorld), C.int32_t(out.N/w.BodiesPerWorld), &out.c))
}

// DownloadColliders — cz_world_download_colliders.
func (w *World) DownloadColliders(firstWorld int, out *HostColliders) {
	check(C.cz_world_download_colliders(w.h, C.int32_t(firstWorld), C.int32_t(out.N/w.BodiesPerWorld), &out.c))
}

// ContactRecord is one generated contact of a world with world-local body indices (-1 = nil).
type ContactRecord struct {
	Body0, Body1                         int32
	Friction, Restitution, Penetration   m.Real
	ContactPoint, ContactNormal          m.Vector3
}

// Contacts — cz_world_download_contacts: the contacts the last frame generated in one world, as generated (before
// ResolveContacts), in the reference's append order.
func (w *World) Contacts(world int) []ContactRecord {
	n := w.ContactsPerWorld
	rs := unsafe.Sizeof(C.cz_real(0))
	var c C.cz_contacts
	c.capacity = C.int32_t(n)
	b0P, b1P := cbuf(n, 4), cbuf(n, 4)
	frP, reP, ptP, nmP, peP := cbuf(n, rs), cbuf(n, rs), cbuf(3*n, rs), cbuf(3*n, rs), cbuf(n, rs)
	for _, p := range []unsafe.Pointer{b0P, b1P, frP, reP, ptP, nmP, peP} {
		defer C.free(p)
	}
	c.body0, c.body1 = (*C.int32_t)(b0P), (*C.int32_t)(b1P)
	c.friction, c.restitution = (*C.cz_real)(frP), (*C.cz_real)(reP)
	c.point, c.normal, c.penetration = (*C.cz_real)(ptP), (*C.cz_real)(nmP), (*C.cz_real)(peP)
	check(C.cz_world_download_contacts(w.h, C.int32_t(world), &c))
	k := int(c.n)
	out := make([]ContactRecord, k)
	b0, b1 := unsafe.Slice((*int32)(b0P), n), unsafe.Slice((*int32)(b1P), n)
	for i := 0; i < k; i++ {
		r := &out[i]
		r.Body0, r.Body1 = b0[i], b1[i]
		r.Friction, r.Restitution, r.Penetration = reals(c.friction, n)[i], reals(c.restitution, n)[i], reals(c.penetration, n)[i]
		copy(r.ContactPoint[:], reals(c.point, 3*n)[3*i:3*i+3])
		copy(r.ContactNormal[:], reals(c.normal, 3*n)[3*i:3*i+3])
	}
	return out
}

// LastStepCounts — cz_world_last_step_counts: contacts, adjustPositions iterations, adjustVelocities iterations of the
// last frame, per world.
func (w *World) LastStepCounts() (contacts, posIterations, velIterations []int32) {
	n := w.NWorlds
	a, b, c := cbuf(n, 4), cbuf(n, 4), cbuf(n, 4)
	defer C.free(a)
	defer C.free(b)
	defer C.free(c)
	check(C.cz_world_last_step_counts(w.h, (*C.int32_t)(a), (*C.int32_t)(b), (*C.int32_t)(c)))
	contacts, posIterations, velIterations = make([]int32, n), make([]int32, n), make([]int32, n)
	copy(contacts, unsafe.Slice((*int32)(a), n))
	copy(posIterations, unsafe.Slice((*int32)(b), n))
	copy(velIterations, unsafe.Slice((*int32)(c), n))
	return
}

// CountNonFinite — cz_world_count_nonfinite: bodies with a NaN / infinite component in their state (the reference, and
// therefore the library, propagate NaN silently: this is how a host notices).
func (w *World) CountNonFinite() int64 {
	var n C.int64_t
	check(C.cz_world_count_nonfinite(w.h, &n))
	return int64(n)
}

// IslandStats — cz_world_island_stats: frames resolved as one CTA per contact island, and how many of those were
// re-run on the single-CTA path because the reference's iteration cap would have cut the loop.
func (w *World) IslandStats() (islandFrames, fallbacks int64) {
	var a, b C.int64_t
	check(C.cz_world_island_stats(w.h, &a, &b))
	return int64(a), int64(b)
}

// ChecksumEnergy — cz_world_checksum_energy (SURVEY §8d definitions).
func (w *World) ChecksumEnergy() (uint64, float64) {
	var c C.uint64_t
	var e C.double
	check(C.cz_world_checksum_energy(w.h, &c, &e))
	return uint64(c), float64(e)
}

// StepHost — cz_world_step_host: the end-to-end call of a host-resident caller.  io (every field) is uploaded, n
// frames run, everything a frame writes is downloaded — chunked, transfers overlapped with the kernels.
func (w *World) StepHost(io *HostBodies, dt m.Real, n int) StepStats {
	var st C.cz_step_stats
	check(C.cz_world_step_host(w.h, &io.c, C.cz_real(dt), C.int32_t(n), &st))
	return statsOf(&st)
}

// StepRL — cz_world_step_rl: the RL loop's frame for device-resident worlds: AddVelocity(addVelocity[3i:]) and
// AddRotation(addRotation[3i:]) on every body (rigidbody.go:195-202; pinned slices from PinnedReals, either may be
// nil), n frames, then the fields obs holds are filled.
func (w *World) StepRL(addVelocity, addRotation []m.Real, obs *HostBodies, dt m.Real, n int) StepStats {
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
	return statsOf(&st)
}

// Obs32 holds float32 observations of the pipelined RL step in pinned memory (PinnedFloats): position 3, orientation 4
// (w, x, y, z), velocity 3, rotation 3 per body — half the download bytes of the Real arrays.
type Obs32 struct {
	c                                         C.cz_obs32
	Position, Orientation, Velocity, Rotation []float32
}

// NewObs32 allocates pinned float32 observation arrays for n bodies.
func NewObs32(n int) *Obs32 {
	o := &Obs32{Position: PinnedFloats(3 * n), Orientation: PinnedFloats(4 * n), Velocity: PinnedFloats(3 * n), Rotation: PinnedFloats(3 * n)}
	o.c.n = C.int32_t(n)
	o.c.position, o.c.orientation = (*C.float)(unsafe.Pointer(&o.Position[0])), (*C.float)(unsafe.Pointer(&o.Orientation[0]))
	o.c.velocity, o.c.rotation = (*C.float)(unsafe.Pointer(&o.Velocity[0])), (*C.float)(unsafe.Pointer(&o.Rotation[0]))
	return o
}

// StepRLAsync — cz_world_step_rl_async: the pipelined form of StepRL.  Returns a ticket at once; up to two steps may be
// in flight (RLWait on the older ticket first), so the observations of step t travel to the host while the frames of
// step t+1 run.  Use two sets of observation buffers and do not touch a set whose step is in flight.
func (w *World) StepRLAsync(addVelocity, addRotation []m.Real, obs *HostBodies, obs32 *Obs32, dt m.Real, n int) int {
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
	var o32 *C.cz_obs32
	if obs32 != nil {
		o32 = &obs32.c
	}
	var ticket C.int32_t
	check(C.cz_world_step_rl_async(w.h, av, ar, o, o32, C.cz_real(dt), C.int32_t(n), &ticket))
	return int(ticket)
}

// RLWait — cz_world_rl_wait: blocks until the observations of the step with this ticket are complete.
func (w *World) RLWait(ticket int) StepStats {
	var st C.cz_step_stats
	check(C.cz_world_rl_wait(w.h, C.int32_t(ticket), &st))
	return statsOf(&st)
}

// ExportGL — cz_world_export_gl: float32 Location (3 per body) and LocalRotation (4 per body: W, V[0], V[1], V[2]) of
// every body of worlds [firstWorld, firstWorld+nWorlds) — the per-frame SetGlVector3 / SetGlQuat copy of
// examples/cubedrop.go:35-37 and examples/exampleapp.go:146-159, converted on the device.  model (optional) receives
// the body transform as a column-major 4x4.  The slices must be pinned or C memory (PinnedFloats).
func (w *World) ExportGL(firstWorld, nWorlds int, location, rotation, model []float32) {
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
	check(C.cz_world_export_gl(w.h, C.int32_t(firstWorld), C.int32_t(nWorlds), l, q, md, 0))
}

// ExportGLDevice is ExportGL into device pointers (a mapped GL buffer), asynchronous on the context stream.
func (w *World) ExportGLDevice(firstWorld, nWorlds int, location, rotation, model unsafe.Pointer) {
	check(C.cz_world_export_gl(w.h, C.int32_t(firstWorld), C.int32_t(nWorlds), (*C.float)(location), (*C.float)(rotation), (*C.float)(model), 1))
}

// PinnedFloats returns a page-locked []float32 (ExportGL destinations).
func PinnedFloats(count int) []float32 {
	var p unsafe.Pointer
	check(C.cz_host_alloc(ctx, C.uint64_t(count)*4, &p))
	return unsafe.Slice((*float32)(p), count)
}
