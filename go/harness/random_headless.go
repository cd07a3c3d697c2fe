// random_headless: the UNMODIFIED reference (github.com/tbogdala/cubez) on the fuzz scene of cubez_b200/scenes.py
// (random_worlds): worlds of random cubes, spheres and collider-less bodies with their own sizes, masses, dampings,
// gravity, spin, sleep flags, activation steps and collider Offset matrices, above one to three half-spaces — the
// branches the five BASELINE configs never reach together.  Same frame loop (examples/cubedrop.go:29-75, every
// collider against every plane, then against every other collider) and same dump as cubedrop_headless.go.
//
//	go run random_headless.go <steps> <worlds> <bodies per world> <seed> <planes> [big 0|1] [material seed] > dump.txt
//
// Also translated by oracle/go2cpp.py into oracle/_ref/random_headless: keep it inside that tool's Go subset.
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

// ---- one world: bodies in creation order; collider k belongs to body owner[k] -------------------------------------
type world struct {
	bodies     []*cubez.RigidBody
	activeFrom []int
	colliders  []cubez.Collider
	owner      []int
	index      map[*cubez.RigidBody]int
	material   []int // per body, when surface materials are on (scenes.with_materials)
}

// Per-pair surface materials as a host of the reference applies them: the contacts a check (one, two) has just
// appended get Friction / Restitution of table[material(one)][material(two)] instead of the test constants 0.9 / 0.1
// (the FIXME at colliders.go:199-202 and its five copies).  The tables are deliberately asymmetric.
func paint(contacts []*cubez.Contact, from int, one int, two int) {
	friction := [9]m.Real{0.9, 0.35, 0.8, 0.3, 0.05, 0.5, 0.8, 0.45, 1.1} // row = material of `one`
	restitution := [9]m.Real{0.1, 0.2, 0.55, 0.25, 0.0, 0.4, 0.6, 0.35, 0.7}
	for k := from; k < len(contacts); k++ {
		contacts[k].Friction = friction[one*3+two]
		contacts[k].Restitution = restitution[one*3+two]
	}
}

// one frame of updateCallback (examples/cubedrop.go:69-75); returns the number of contacts generated
func (w *world) step(planes []*cubez.CollisionPlane, step int, dt m.Real, pairs *uint64, gen *hasher) int {
	for i, b := range w.bodies { // updateObjects, cubedrop.go:29-39 (a body without a collider is integrated too)
		if step < w.activeFrom[i] {
			continue
		}
		b.Integrate(dt)
	}
	for k, c := range w.colliders {
		if step < w.activeFrom[w.owner[k]] {
			continue
		}
		c.CalculateDerivedData()
	}
	var contacts []*cubez.Contact // generateContacts, cubedrop.go:42-67
	found := false
	for k, c := range w.colliders {
		if step < w.activeFrom[w.owner[k]] {
			continue
		}
		var f bool
		for pi, p := range planes {
			before := len(contacts)
			f, contacts = c.CheckAgainstHalfSpace(p, contacts)
			if f {
				found = true
			}
			if len(w.material) > 0 {
				planeMaterial := 0
				if pi == 0 {
					planeMaterial = 1 // the first plane is slippery
				}
				paint(contacts, before, w.material[w.owner[k]], planeMaterial)
			}
		}
		for j, o := range w.colliders {
			if j == k || step < w.activeFrom[w.owner[j]] {
				continue
			}
			before := len(contacts)
			f, contacts = cubez.CheckForCollisions(c, o, contacts)
			if f {
				found = true
			}
			if len(w.material) > 0 {
				paint(contacts, before, w.material[w.owner[k]], w.material[w.owner[j]])
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

// scenes.random_worlds: body g (global index) draws 40 uniforms from splitmix64(seed*0x100000001B3 + g*0x9E3779B97F4A7C15)
func build(w *world, first int, count int, seed uint64, extent float64, height float64, matSeed int) {
	for i := 0; i < count; i++ {
		g := uint64(first + i)
		if matSeed > 0 { // scenes.with_materials: id = min(int(U(0,3)), 2) from splitmix64(matSeed + g)
			ms := uint64(matSeed) + g
			id := int(uniform(draw(&ms), 0.0, 3.0))
			if id > 2 {
				id = 2
			}
			w.material = append(w.material, id)
		}
		state := seed*0x100000001B3 + g*0x9E3779B97F4A7C15
		var u [40]float64
		for k := 0; k < 40; k++ {
			u[k] = draw(&state)
		}
		kind := uniform(u[0], 0.0, 1.0)
		half := m.Vector3{m.Real(uniform(u[1], 0.25, 0.75)), m.Real(uniform(u[2], 0.25, 0.75)), m.Real(uniform(u[3], 0.25, 0.75))}
		radius := m.Real(uniform(u[4], 0.25, 0.7))
		mass := m.Real(uniform(u[5], 1.0, 10.0))
		var body *cubez.RigidBody
		var cube *cubez.CollisionCube
		var sphere *cubez.CollisionSphere
		var inertia m.Matrix3
		if kind < 0.47 {
			cube = cubez.NewCollisionCube(nil, half)
			body = cube.Body
			body.SetMass(mass)
			inertia.SetBlockInertiaTensor(&cube.HalfSize, mass)
		} else {
			var r m.Real = 0.5
			if kind < 0.92 {
				sphere = cubez.NewCollisionSphere(nil, radius)
				body = sphere.Body
				r = radius
			} else {
				body = cubez.NewRigidBody()
			}
			body.SetMass(mass)
			var coeff m.Real = 0.4 * mass * r * r
			inertia.SetInertiaTensorCoeffs(coeff, coeff, coeff, 0.0, 0.0, 0.0)
		}
		body.SetInertiaTensor(&inertia)
		body.Position = m.Vector3{m.Real(uniform(u[6], -extent, extent)), m.Real(uniform(u[8], 0.9, height)), m.Real(uniform(u[7], -extent, extent))}
		q0 := uniform(u[9], -1.0, 1.0)
		q1 := uniform(u[10], -1.0, 1.0)
		q2 := uniform(u[11], -1.0, 1.0)
		q3 := uniform(u[12], -1.0, 1.0)
		if math.Sqrt(q0*q0+q1*q1+q2*q2+q3*q3) < 0.1 {
			q0 = 1.0
			q1 = 0.0
			q2 = 0.0
			q3 = 0.0
		}
		r0 := m.Real(q0)
		r1 := m.Real(q1)
		r2 := m.Real(q2)
		r3 := m.Real(q3)
		ln := m.Real(math.Sqrt(float64(((r0*r0 + r1*r1) + r2*r2) + r3*r3)))
		body.Orientation = m.Quat{r0 / ln, r1 / ln, r2 / ln, r3 / ln}
		body.Velocity = m.Vector3{m.Real(uniform(u[13], -2.0, 2.0)), m.Real(uniform(u[14], -2.0, 2.0)), m.Real(uniform(u[15], -2.0, 2.0))}
		body.Rotation = m.Vector3{m.Real(uniform(u[16], -2.0, 2.0)), m.Real(uniform(u[17], -2.0, 2.0)), m.Real(uniform(u[18], -2.0, 2.0))}
		body.LinearDamping = m.Real(uniform(u[19], 0.90, 0.99))
		body.AngularDamping = m.Real(uniform(u[20], 0.85, 0.99))
		if uniform(u[21], 0.0, 1.0) < 0.1 {
			body.Acceleration = m.Vector3{0.0, 0.0, 0.0}
		}
		body.CanSleep = uniform(u[22], 0.0, 1.0) < 0.8
		if uniform(u[23], 0.0, 1.0) < 0.1 {
			body.SetAwake(false)
		}
		from := 0
		if uniform(u[29], 0.0, 1.0) < 0.3 {
			from = int(uniform(u[30], 1.0, 40.0))
		}
		body.CalculateDerivedData()
		w.index[body] = len(w.bodies)
		w.bodies = append(w.bodies, body)
		w.activeFrom = append(w.activeFrom, from)
		if cube == nil && sphere == nil {
			continue
		}
		// collider Offset: rotation about z from a table of Pythagorean (cos, sin) pairs, and a small translation
		var offset m.Matrix3x4
		offset.SetIdentity()
		if uniform(u[24], 0.0, 1.0) < 0.33 {
			cosT := [6]m.Real{0.8, 0.6, 0.96, 0.8, 0.6, 0.96}
			sinT := [6]m.Real{0.6, 0.8, 0.28, -0.6, -0.8, -0.28}
			pick := int(uniform(u[25], 0.0, 6.0))
			if pick > 5 {
				pick = 5
			}
			offset[0] = cosT[pick]
			offset[1] = sinT[pick]
			offset[3] = -sinT[pick]
			offset[4] = cosT[pick]
			offset[9] = m.Real(uniform(u[26], -0.2, 0.2))
			offset[10] = m.Real(uniform(u[27], -0.2, 0.2))
			offset[11] = m.Real(uniform(u[28], -0.2, 0.2))
		}
		if cube != nil {
			cube.Offset = offset
			cube.CalculateDerivedData()
			w.colliders = append(w.colliders, cube)
		} else {
			sphere.Offset = offset
			sphere.CalculateDerivedData()
			w.colliders = append(w.colliders, sphere)
		}
		w.owner = append(w.owner, len(w.bodies)-1)
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
	steps := argInt(1, 120)
	nWorlds := argInt(2, 8)
	perWorld := argInt(3, 8)
	seed := argInt(4, 11)
	nPlanes := argInt(5, 2)
	extent := 1.8
	height := 6.0
	if argInt(6, 0) > 0 { // one large world: spread out
		extent = 5.0
		height = 9.0
	}
	matSeed := argInt(7, 0) // > 0: per-pair surface materials with ids drawn from this seed
	planes := []*cubez.CollisionPlane{cubez.NewCollisionPlane(m.Vector3{0.0, 1.0, 0.0}, 0.0)}
	if nPlanes >= 2 {
		planes = append(planes, cubez.NewCollisionPlane(m.Vector3{1.0, 0.0, 0.0}, m.Real(-(extent + 1.2))))
	}
	if nPlanes >= 3 {
		planes = append(planes, cubez.NewCollisionPlane(m.Vector3{-0.6, 0.8, 0.0}, -1.5))
	}
	var worlds []*world
	for k := 0; k < nWorlds; k++ {
		w := new(world)
		w.index = make(map[*cubez.RigidBody]int)
		build(w, k*perWorld, perWorld, uint64(seed), extent, height, matSeed)
		worlds = append(worlds, w)
	}
	dt := m.Real(1.0 / 60.0)
	fmt.Printf("cubez-dump v2 scene=random_worlds worlds=%d bodies=%d steps=%d seed=%d planes=%d big=%d materials=%d\n", nWorlds, nWorlds*perWorld, steps, seed, nPlanes, argInt(6, 0), matSeed)
	start := time.Now()
	for s := 0; s < steps; s++ {
		var pairs uint64
		var gen hasher
		var state hasher
		total := 0
		for _, w := range worlds {
			total += w.step(planes, s, dt, &pairs, &gen)
		}
		for _, w := range worlds {
			for _, b := range w.bodies {
				state.body(b)
			}
		}
		fmt.Printf("step %d contacts %d pairhash %016x genhash %016x statehash %016x\n", s, total, pairs, gen.h, state.h)
	}
	wall := time.Since(start).Seconds()
	n := 0
	for _, w := range worlds {
		for _, b := range w.bodies {
			if n < 64 {
				printBody(n, b)
			}
			n++
		}
	}
	fmt.Fprintf(os.Stderr, "wall %.6f s for %d steps of %d worlds\n", wall, steps, nWorlds)
}
