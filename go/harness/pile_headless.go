// pile_headless: the UNMODIFIED reference (github.com/tbogdala/cubez) on BASELINE config 3 — a side^3 jittered lattice of
// alternating cubes and spheres falling into a pile on the ground plane, stepped by the all-pairs-ordered loop of
// examples/cubedrop.go:29-75 (16.8 M ordered checks per frame at side = 16: this is the reference's O(n^2) stress).
// Same dump as cubedrop_headless.go.
//
//	go run pile_headless.go <steps> [side] > dump.txt ; python tools/compare_go_dump.py dump.txt
//
// Also translated by oracle/go2cpp.py into oracle/_ref/pile_headless: keep it inside that tool's Go subset.
package main

import (
	"fmt"
	"math"
	"os"
	"strconv"
	"time"

	"github.com/tbogdala/cubez"
	m "github.com/tbogdala/cubez/math"
)

// ---- dump helpers (identical in every harness) ---------------------------------------------------------------
func mix(z uint64) uint64 { // splitmix64 finaliser
	z += 0x9E3779B97F4A7C15
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9
	z = (z ^ (z >> 27)) * 0x94D049BB133111EB
	return z ^ (z >> 31)
}

// hasher: h = sum over elements k of mix(bits_k ^ mix(k)) mod 2^64 (position-salted, so it vectorises in numpy)
type hasher struct {
	h uint64
	k uint64
}

func (s *hasher) add(bits uint64) {
	s.h += mix(bits ^ mix(s.k))
	s.k++
}
func (s *hasher) real(r m.Real) { s.add(math.Float64bits(float64(r))) }
func (s *hasher) vec3(v *m.Vector3) {
	s.real(v[0])
	s.real(v[1])
	s.real(v[2])
}
func (s *hasher) body(b *cubez.RigidBody) {
	s.vec3(&b.Position)
	s.real(b.Orientation[0])
	s.real(b.Orientation[1])
	s.real(b.Orientation[2])
	s.real(b.Orientation[3])
	s.vec3(&b.Velocity)
	s.vec3(&b.Rotation)
	if b.IsAwake {
		s.add(1)
	} else {
		s.add(0)
	}
}
func bits(r m.Real) uint64 { return math.Float64bits(float64(r)) }

func draw(state *uint64) float64 { // splitmix64 -> [0,1)
	*state += 0x9E3779B97F4A7C15
	z := *state
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9
	z = (z ^ (z >> 27)) * 0x94D049BB133111EB
	z = z ^ (z >> 31)
	return float64(z>>11) * (1.0 / 9007199254740992.0)
}
func uniform(u float64, a float64, b float64) float64 { return a + (b-a)*u }

func printBody(i int, b *cubez.RigidBody) {
	awake := 0
	if b.IsAwake {
		awake = 1
	}
	fmt.Printf("body %d %016x %016x %016x %016x %016x %016x %016x %016x %016x %016x %016x %016x %016x awake %d\n", i,
		bits(b.Position[0]), bits(b.Position[1]), bits(b.Position[2]),
		bits(b.Orientation[0]), bits(b.Orientation[1]), bits(b.Orientation[2]), bits(b.Orientation[3]),
		bits(b.Velocity[0]), bits(b.Velocity[1]), bits(b.Velocity[2]),
		bits(b.Rotation[0]), bits(b.Rotation[1]), bits(b.Rotation[2]), awake)
}

// ---- one world: colliders in creation order, all-pairs-ordered schedule (examples/cubedrop.go:42-67) ---------------
type world struct {
	colliders  []cubez.Collider
	activeFrom []int
	index      map[*cubez.RigidBody]int
}

func (w *world) add(c cubez.Collider, from int) {
	w.index[c.GetBody()] = len(w.colliders)
	w.colliders = append(w.colliders, c)
	w.activeFrom = append(w.activeFrom, from)
}

// one frame of updateCallback (examples/cubedrop.go:69-75); returns the number of contacts generated
func (w *world) step(ground *cubez.CollisionPlane, step int, dt m.Real, pairs *uint64, gen *hasher) int {
	for i, c := range w.colliders { // updateObjects, cubedrop.go:29-39
		if step < w.activeFrom[i] {
			continue
		}
		c.GetBody().Integrate(dt)
		c.CalculateDerivedData()
	}
	var contacts []*cubez.Contact // generateContacts, cubedrop.go:42-67
	found := false
	for i, c := range w.colliders {
		if step < w.activeFrom[i] {
			continue
		}
		var f bool
		f, contacts = c.CheckAgainstHalfSpace(ground, contacts)
		if f {
			found = true
		}
		for j, o := range w.colliders {
			if j == i || step < w.activeFrom[j] {
				continue
			}
			f, contacts = cubez.CheckForCollisions(c, o, contacts)
			if f {
				found = true
			}
		}
	}
	h := uint64(0xcbf29ce484222325) // FNV-1a over the (body0, body1) index sequence, nil = 0xFFFFFFFF
	for _, c := range contacts {
		for k := 0; k < 2; k++ {
			v := uint64(0xFFFFFFFF)
			if c.Bodies[k] != nil {
				v = uint64(w.index[c.Bodies[k]])
			}
			h = (h ^ v) * 0x100000001b3
		}
		gen.vec3(&c.ContactPoint)
		gen.vec3(&c.ContactNormal)
		gen.real(c.Penetration)
	}
	*pairs += h
	if found {
		cubez.ResolveContacts(len(contacts)*8, contacts, dt) // cubedrop.go:72-74
	}
	return len(contacts)
}

// SURVEY section 8d, cfg3: body b = ix + side*(iz + side*iy) at (1.25*(ix-h)+jx, 0.75+1.25*iy+jy, 1.25*(iz-h)+jz),
// jitter U(-0.05, 0.05) from splitmix64(0xC0BE2 + b); (ix+iy+iz) even: cube half 0.5 mass 8, odd: sphere r 0.5 mass 4
func build(w *world, side int) {
	h := float64(side-1) / 2.0
	for b := 0; b < side*side*side; b++ {
		ix := b % side
		iz := (b / side) % side
		iy := b / (side * side)
		state := uint64(0xC0BE2) + uint64(b)
		jx := uniform(draw(&state), -0.05, 0.05)
		jy := uniform(draw(&state), -0.05, 0.05)
		jz := uniform(draw(&state), -0.05, 0.05)
		pos := m.Vector3{m.Real(1.25*(float64(ix)-h) + jx), m.Real(0.75 + 1.25*float64(iy) + jy), m.Real(1.25*(float64(iz)-h) + jz)}
		var inertia m.Matrix3
		if (ix+iy+iz)%2 == 0 {
			c := cubez.NewCollisionCube(nil, m.Vector3{0.5, 0.5, 0.5})
			c.Body.Position = pos
			c.Body.SetMass(8.0)
			inertia.SetBlockInertiaTensor(&c.HalfSize, 8.0)
			c.Body.SetInertiaTensor(&inertia)
			c.Body.CalculateDerivedData()
			c.CalculateDerivedData()
			w.add(c, 0)
		} else {
			var mass m.Real = 4.0
			var radius m.Real = 0.5
			s := cubez.NewCollisionSphere(nil, radius)
			s.Body.Position = pos
			s.Body.SetMass(mass)
			var coeff m.Real = 0.4 * mass * radius * radius
			inertia.SetInertiaTensorCoeffs(coeff, coeff, coeff, 0.0, 0.0, 0.0)
			s.Body.SetInertiaTensor(&inertia)
			s.Body.CalculateDerivedData()
			s.CalculateDerivedData()
			w.add(s, 0)
		}
	}
}

func argInt(k int, dflt int) int {
	if len(os.Args) > k {
		v, err := strconv.Atoi(os.Args[k])
		if err == nil {
			return v
		}
	}
	return dflt
}

func main() {
	steps := argInt(1, 60)
	side := argInt(2, 16)
	ground := cubez.NewCollisionPlane(m.Vector3{0.0, 1.0, 0.0}, 0.0)
	w := new(world)
	w.index = make(map[*cubez.RigidBody]int)
	build(w, side)
	dt := m.Real(1.0 / 60.0)
	fmt.Printf("cubez-dump v2 scene=pile worlds=1 bodies=%d steps=%d\n", len(w.colliders), steps)
	start := time.Now()
	for s := 0; s < steps; s++ {
		var pairs uint64
		var gen hasher
		var state hasher
		total := w.step(ground, s, dt, &pairs, &gen)
		for _, c := range w.colliders {
			state.body(c.GetBody())
		}
		fmt.Printf("step %d contacts %d pairhash %016x genhash %016x statehash %016x\n", s, total, pairs, gen.h, state.h)
		fmt.Fprintf(os.Stderr, "frame %d: %d contacts, %.3f s so far\n", s, total, time.Since(start).Seconds())
	}
	wall := time.Since(start).Seconds()
	for i, c := range w.colliders {
		if i < 64 {
			printBody(i, c.GetBody())
		}
	}
	fmt.Fprintf(os.Stderr, "wall %.6f s for %d steps of %d bodies\n", wall, steps, len(w.colliders))
}
