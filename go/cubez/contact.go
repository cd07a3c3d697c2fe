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

// Contact — contact.go:17-51 (public fields; the private work fields live on the device during ResolveContacts).
type Contact struct {
	Bodies                      [2]*RigidBody
	Friction, Restitution       m.Real
	ContactPoint, ContactNormal m.Vector3
	Penetration                 m.Real
}

// NewContact — contact.go:54-57.
func NewContact() *Contact { return new(Contact) }

// ResolveContacts — contact.go:208-222: prepareContacts, adjustPositions(maxIterations), adjustVelocities(maxIterations)
// on the device (cz_resolve_contacts).  Mirrors back everything the reference's call visibly mutates (SURVEY §8b): each
// touched body's Position, Orientation, Velocity, Rotation, IsAwake, motion (and transform + world inertia for bodies
// that were asleep, contact.go:380-382); each contact's Penetration, and — where Bodies[0] was nil — the swapped Bodies
// and the negated ContactNormal (contact.go:61-65).  A frictionless one-body contact panics, as the reference does
// (nil dereference at contact.go:512-523 -> CZ_ERR_NIL_BODY).
func ResolveContacts(maxIterations int, contacts []*Contact, duration m.Real) {
	if duration <= 0.0 || contacts == nil || len(contacts) == 0 { // contact.go:210-212
		return
	}
	var bodies []*RigidBody
	bidx := map[*RigidBody]int{}
	for _, c := range contacts {
		for _, b := range c.Bodies {
			if b != nil {
				if _, have := bidx[b]; !have {
					bidx[b] = len(bodies)
					bodies = append(bodies, b)
				}
			}
		}
	}
	if len(bodies) == 0 {
		panic("cubez: ResolveContacts: a contact without bodies (the reference dereferences nil, contact.go:66-70)")
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	n := len(contacts)
	rs := unsafe.Sizeof(C.cz_real(0))
	var cs C.cz_contacts
	cs.capacity, cs.n = C.int32_t(n), C.int32_t(n)
	b0P, b1P := cbuf(n, 4), cbuf(n, 4)
	frP, reP, ptP, nmP, peP := cbuf(n, rs), cbuf(n, rs), cbuf(n*3, rs), cbuf(n*3, rs), cbuf(n, rs)
	for _, p := range []unsafe.Pointer{b0P, b1P, frP, reP, ptP, nmP, peP} {
		defer C.free(p)
	}
	cs.body0, cs.body1 = (*C.int32_t)(b0P), (*C.int32_t)(b1P)
	cs.friction, cs.restitution = (*C.cz_real)(frP), (*C.cz_real)(reP)
	cs.point, cs.normal, cs.penetration = (*C.cz_real)(ptP), (*C.cz_real)(nmP), (*C.cz_real)(peP)
	b0, b1 := unsafe.Slice((*int32)(b0P), n), unsafe.Slice((*int32)(b1P), n)
	fr, re, pe := reals(cs.friction, n), reals(cs.restitution, n), reals(cs.penetration, n)
	pt, nm := reals(cs.point, 3*n), reals(cs.normal, 3*n)
	for i, c := range contacts {
		b0[i], b1[i] = -1, -1
		if c.Bodies[0] != nil {
			b0[i] = int32(bidx[c.Bodies[0]])
		}
		if c.Bodies[1] != nil {
			b1[i] = int32(bidx[c.Bodies[1]])
		}
		fr[i], re[i], pe[i] = c.Friction, c.Restitution, c.Penetration
		copy(pt[3*i:], c.ContactPoint[:])
		copy(nm[3*i:], c.ContactNormal[:])
	}
	f := gather(bodies)
	defer f.free()
	check(C.cz_resolve_contacts(ctx, C.int32_t(maxIterations), &cs, &f.c, C.cz_real(duration), nil))
	f.scatter(bodies)
	for i, c := range contacts {
		c.Bodies[0], c.Bodies[1] = nil, nil
		if b0[i] >= 0 {
			c.Bodies[0] = bodies[b0[i]]
		}
		if b1[i] >= 0 {
			c.Bodies[1] = bodies[b1[i]]
		}
		copy(c.ContactNormal[:], nm[3*i:3*i+3])
		c.Penetration = pe[i]
	}
}
