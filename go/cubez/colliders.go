package cubez

/*
#include <stdlib.h>
#include "cubezcuda.h"
*/
import "C"

import (
	"runtime"
	"unsafe"

	m "github.com/tbogdala/cubez/math"
)

// ---------------------------------------------------------------------------------------------
// Colliders — colliders.go:15-71.  Every CheckAgainst* is CheckForCollisions with the operands in the order the
// reference's method would test them (cz_narrowphase reproduces the dispatch of colliders.go:720-747).
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

// CollisionPlane — colliders.go:29-35.
type CollisionPlane struct {
	Normal m.Vector3
	Offset m.Real
}

// CollisionCube — colliders.go:39-53.
type CollisionCube struct {
	Body      *RigidBody
	Offset    m.Matrix3x4
	transform m.Matrix3x4
	HalfSize  m.Vector3
}

// CollisionSphere — colliders.go:57-71.
type CollisionSphere struct {
	Body      *RigidBody
	Offset    m.Matrix3x4
	transform m.Matrix3x4
	Radius    m.Real
}

func NewCollisionPlane(n m.Vector3, o m.Real) *CollisionPlane { return &CollisionPlane{n, o} }
func NewCollisionCube(optBody *RigidBody, halfSize m.Vector3) *CollisionCube { // colliders.go:265-275
	c := &CollisionCube{Body: optBody, HalfSize: halfSize}
	c.Offset.SetIdentity()
	if c.Body == nil {
		c.Body = NewRigidBody()
	}
	return c
}
func NewCollisionSphere(optBody *RigidBody, radius m.Real) *CollisionSphere { // colliders.go:136-146
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
func (s *CollisionSphere) GetBody() *RigidBody       { return s.Body }
func (s *CollisionSphere) GetTransform() m.Matrix3x4 { return s.transform }
func (c *CollisionCube) Clone() Collider { // colliders.go:277-286
	var body *RigidBody
	if c.Body != nil {
		body = c.Body.Clone()
	}
	n := NewCollisionCube(body, c.HalfSize)
	n.Offset, n.transform = c.Offset, c.transform
	return n
}
func (s *CollisionSphere) Clone() Collider { // colliders.go:148-157
	var body *RigidBody
	if s.Body != nil {
		body = s.Body.Clone()
	}
	n := NewCollisionSphere(body, s.Radius)
	n.Offset, n.transform = s.Offset, s.transform
	return n
}

// derive: transform = body.transform x Offset (colliders.go:173-176 / 302-304) through C memory
func derive(body *RigidBody, offset *m.Matrix3x4) (out m.Matrix3x4) {
	buf := (*C.cz_real)(C.malloc(C.size_t(36) * C.size_t(unsafe.Sizeof(C.cz_real(0)))))
	defer C.free(unsafe.Pointer(buf))
	r := reals(buf, 36)
	copy(r[0:12], body.transform[:])
	copy(r[12:24], offset[:])
	tr := (*C.cz_real)(unsafe.Pointer(&r[0]))
	of := (*C.cz_real)(unsafe.Pointer(&r[12]))
	res := (*C.cz_real)(unsafe.Pointer(&r[24]))
	check(C.cz_collider_derive(ctx, 1, tr, of, res))
	copy(out[:], r[24:36])
	return
}
func (c *CollisionCube) CalculateDerivedData()   { c.transform = derive(c.Body, &c.Offset) }
func (s *CollisionSphere) CalculateDerivedData() { s.transform = derive(s.Body, &s.Offset) }

func (p *CollisionPlane) CheckAgainstHalfSpace(_ *CollisionPlane, e []*Contact) (bool, []*Contact) {
	return false, e // colliders.go:111-113
}
func (p *CollisionPlane) CheckAgainstSphere(s *CollisionSphere, e []*Contact) (bool, []*Contact) {
	return CheckForCollisions(p, s, e)
}
func (p *CollisionPlane) CheckAgainstCube(c *CollisionCube, e []*Contact) (bool, []*Contact) {
	return CheckForCollisions(p, c, e)
}
func (c *CollisionCube) CheckAgainstHalfSpace(p *CollisionPlane, e []*Contact) (bool, []*Contact) {
	return CheckForCollisions(c, p, e)
}
func (c *CollisionCube) CheckAgainstSphere(s *CollisionSphere, e []*Contact) (bool, []*Contact) {
	return CheckForCollisions(c, s, e)
}
func (c *CollisionCube) CheckAgainstCube(o *CollisionCube, e []*Contact) (bool, []*Contact) {
	return CheckForCollisions(c, o, e)
}
func (s *CollisionSphere) CheckAgainstHalfSpace(p *CollisionPlane, e []*Contact) (bool, []*Contact) {
	return CheckForCollisions(s, p, e)
}
func (s *CollisionSphere) CheckAgainstSphere(o *CollisionSphere, e []*Contact) (bool, []*Contact) {
	return CheckForCollisions(s, o, e)
}
func (s *CollisionSphere) CheckAgainstCube(c *CollisionCube, e []*Contact) (bool, []*Contact) {
	return CheckForCollisions(s, c, e)
}

// Check is one ordered (one, two) entry of a pair schedule.
type Check struct{ One, Two Collider }

// CheckForCollisions — colliders.go:720-747.
func CheckForCollisions(one Collider, two Collider, existingContacts []*Contact) (bool, []*Contact) {
	found, contacts := CheckCollisionList([]Check{{one, two}}, existingContacts)
	return found[0], contacts
}

// cbuf is a block of C memory holding `n` elements of `size` bytes (freed by the caller).
func cbuf(n int, size uintptr) unsafe.Pointer {
	if n < 1 {
		n = 1
	}
	return C.malloc(C.size_t(n) * C.size_t(size))
}

// CheckCollisionList evaluates an ordered list of checks in ONE library call (cz_narrowphase); contacts are appended
// in the order the reference's append calls would produce (check order, then vertex order).  The pointer graph is
// flattened to indices first — cgo forbids Go pointers inside C memory: colliders, planes and bodies are numbered in
// first-use order; the returned body indices are mapped back to *RigidBody.
func CheckCollisionList(checks []Check, existing []*Contact) ([]bool, []*Contact) {
	contacts := existing
	found := make([]bool, len(checks))
	if len(checks) == 0 {
		return found, contacts
	}
	var colliders []Collider
	var planes []*CollisionPlane
	var bodies []*RigidBody
	cidx := map[Collider]int{}
	pidx := map[*CollisionPlane]int{}
	bidx := map[*RigidBody]int{}
	reg := func(x Collider) int32 {
		if p, ok := x.(*CollisionPlane); ok {
			k, seen := pidx[p]
			if !seen {
				k = len(planes)
				pidx[p] = k
				planes = append(planes, p)
			}
			return int32(-(k + 1))
		}
		k, seen := cidx[x]
		if !seen {
			k = len(colliders)
			cidx[x] = k
			colliders = append(colliders, x)
			if b := x.GetBody(); b != nil {
				if _, have := bidx[b]; !have {
					bidx[b] = len(bodies)
					bodies = append(bodies, b)
				}
			}
		}
		return int32(k)
	}
	nChk := len(checks)
	oneP := cbuf(nChk, 4)
	twoP := cbuf(nChk, 4)
	defer C.free(oneP)
	defer C.free(twoP)
	one, two := unsafe.Slice((*int32)(oneP), nChk), unsafe.Slice((*int32)(twoP), nChk)
	for k, c := range checks {
		one[k], two[k] = reg(c.One), reg(c.Two)
	}
	if len(colliders) == 0 { // plane against plane only: colliders.go:111-113
		return found, contacts
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()

	// colliders -> cz_colliders
	nc := len(colliders)
	rs := unsafe.Sizeof(C.cz_real(0))
	var cc C.cz_colliders
	cc.n = C.int32_t(nc)
	shapeP, bodyP := cbuf(nc, 4), cbuf(nc, 4)
	offP, trP, halfP, radP := cbuf(nc*12, rs), cbuf(nc*12, rs), cbuf(nc*3, rs), cbuf(nc, rs)
	for _, p := range []unsafe.Pointer{shapeP, bodyP, offP, trP, halfP, radP} {
		defer C.free(p)
	}
	cc.shape, cc.body = (*C.int32_t)(shapeP), (*C.int32_t)(bodyP)
	cc.offset, cc.transform, cc.half_size, cc.radius = (*C.cz_real)(offP), (*C.cz_real)(trP), (*C.cz_real)(halfP), (*C.cz_real)(radP)
	shape, body := unsafe.Slice((*int32)(shapeP), nc), unsafe.Slice((*int32)(bodyP), nc)
	off, tr, half, rad := reals(cc.offset, 12*nc), reals(cc.transform, 12*nc), reals(cc.half_size, 3*nc), reals(cc.radius, nc)
	for i, c := range colliders {
		switch v := c.(type) {
		case *CollisionCube:
			shape[i] = C.CZ_SHAPE_CUBE
			copy(off[12*i:], v.Offset[:])
			copy(tr[12*i:], v.transform[:])
			copy(half[3*i:], v.HalfSize[:])
			rad[i] = 0
		case *CollisionSphere:
			shape[i] = C.CZ_SHAPE_SPHERE
			copy(off[12*i:], v.Offset[:])
			copy(tr[12*i:], v.transform[:])
			half[3*i], half[3*i+1], half[3*i+2] = 0, 0, 0
			rad[i] = v.Radius
		default:
			panic("cubez: unsupported Collider implementation")
		}
		body[i] = int32(bidx[c.GetBody()])
	}
	// planes -> cz_planes
	var cp C.cz_planes
	var planesArg *C.cz_planes
	if len(planes) > 0 {
		np := len(planes)
		nP, oP := cbuf(np*3, rs), cbuf(np, rs)
		defer C.free(nP)
		defer C.free(oP)
		cp.n, cp.normal, cp.offset = C.int32_t(np), (*C.cz_real)(nP), (*C.cz_real)(oP)
		for i, p := range planes {
			copy(reals(cp.normal, 3*np)[3*i:], p.Normal[:])
			reals(cp.offset, np)[i] = p.Offset
		}
		planesArg = &cp
	}
	// bodies: cz_narrowphase reads Velocity only (cube-sphere fallback normal, colliders.go:417-421)
	f := gather(bodies)
	defer f.free()

	// output: 8 contacts per check is the maximum (cube against a plane); grow and retry on CZ_ERR_CAPACITY anyway
	capacity := 8 * nChk
	foundP := cbuf(nChk, 1)
	defer C.free(foundP)
	for {
		var out C.cz_contacts
		out.capacity = C.int32_t(capacity)
		b0P, b1P := cbuf(capacity, 4), cbuf(capacity, 4)
		frP, reP, ptP, nmP, peP := cbuf(capacity, rs), cbuf(capacity, rs), cbuf(capacity*3, rs), cbuf(capacity*3, rs), cbuf(capacity, rs)
		out.body0, out.body1 = (*C.int32_t)(b0P), (*C.int32_t)(b1P)
		out.friction, out.restitution = (*C.cz_real)(frP), (*C.cz_real)(reP)
		out.point, out.normal, out.penetration = (*C.cz_real)(ptP), (*C.cz_real)(nmP), (*C.cz_real)(peP)
		rc := C.cz_narrowphase(ctx, &cc, planesArg, &f.c, C.int32_t(nChk), (*C.int32_t)(oneP), (*C.int32_t)(twoP), &out, (*C.uint8_t)(foundP))
		n := int(out.n)
		if rc == C.CZ_OK {
			b0, b1 := unsafe.Slice((*int32)(b0P), capacity), unsafe.Slice((*int32)(b1P), capacity)
			fr, re, pe := reals(out.friction, capacity), reals(out.restitution, capacity), reals(out.penetration, capacity)
			pt, nm := reals(out.point, 3*capacity), reals(out.normal, 3*capacity)
			for k := 0; k < n; k++ {
				c := NewContact()
				if b0[k] >= 0 {
					c.Bodies[0] = bodies[b0[k]]
				}
				if b1[k] >= 0 {
					c.Bodies[1] = bodies[b1[k]]
				}
				c.Friction, c.Restitution, c.Penetration = fr[k], re[k], pe[k]
				copy(c.ContactPoint[:], pt[3*k:3*k+3])
				copy(c.ContactNormal[:], nm[3*k:3*k+3])
				contacts = append(contacts, c)
			}
			fnd := unsafe.Slice((*uint8)(foundP), nChk)
			for k := range found {
				found[k] = fnd[k] != 0
			}
		}
		for _, p := range []unsafe.Pointer{b0P, b1P, frP, reP, ptP, nmP, peP} {
			C.free(p)
		}
		if rc == C.CZ_ERR_CAPACITY && n > capacity {
			capacity = n
			continue
		}
		check(rc)
		return found, contacts
	}
}
