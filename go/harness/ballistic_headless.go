// ballistic_headless: the UNMODIFIED reference (github.com/tbogdala/cubez) on BASELINE config 2 — examples/ballistic.go
// without GL: a cube resting on the ground plane, a static backboard, and bullets fired at them (bullet k is created at
// the start of frame first + every*k, SURVEY section 8d).  Same dump as cubedrop_headless.go.
//
//	go run ballistic_headless.go <steps> [bullets [first [every]]] > dump.txt ; python tools/compare_go_dump.py dump.txt
//
// Also translated by oracle/go2cpp.py into oracle/_ref/ballistic_headless: keep it inside that tool's Go subset.
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

// ---- the scene of examples/ballistic.go:149-231 -------------------------------------------------------------
var cube *cubez.CollisionCube
var backboard *cubez.CollisionCube
var bullets []*cubez.CollisionSphere
var index map[*cubez.RigidBody]int

func setup() { // examples/ballistic.go:163-187
	var cubeMass m.Real = 8.0
	var cubeInertia m.Matrix3
	cube = cubez.NewCollisionCube(nil, m.Vector3{1.0, 1.0, 1.0})
	cube.Body.Position = m.Vector3{0.0, 5.0, 0.0}
	cube.Body.SetMass(cubeMass)
	cubeInertia.SetBlockInertiaTensor(&cube.HalfSize, cubeMass)
	cube.Body.SetInertiaTensor(&cubeInertia)
	cube.Body.CalculateDerivedData()
	cube.CalculateDerivedData()
	backboard = cubez.NewCollisionCube(nil, m.Vector3{0.5, 2.0, 0.25})
	backboard.Body.Position = m.Vector3{0.0, 2.0, -10.0}
	backboard.Body.SetInfiniteMass()
	backboard.Body.CalculateDerivedData()
	backboard.CalculateDerivedData()
	index = make(map[*cubez.RigidBody]int)
	index[cube.Body] = 0
	index[backboard.Body] = 1
}

func fire() { // examples/ballistic.go:204-231
	var mass m.Real = 1.5
	var radius m.Real = 0.2
	bullet := cubez.NewCollisionSphere(nil, radius)
	bullet.Body.Position = m.Vector3{0.0, 1.5, 20.0}
	var inertia m.Matrix3
	var coeff m.Real = 0.4 * mass * radius * radius
	inertia.SetInertiaTensorCoeffs(coeff, coeff, coeff, 0.0, 0.0, 0.0)
	bullet.GetBody().SetInertiaTensor(&inertia)
	bullet.Body.SetMass(mass)
	bullet.Body.Velocity = m.Vector3{0.0, 0.0, -40.0}
	bullet.Body.Acceleration = m.Vector3{0.0, -2.5, 0.0}
	bullet.Body.CalculateDerivedData()
	bullet.CalculateDerivedData()
	index[bullet.Body] = 2 + len(bullets)
	bullets = append(bullets, bullet)
}

// one frame of updateCallback (examples/ballistic.go:99-105)
func step(dt m.Real, pairs *uint64, gen *hasher) int {
	cube.Body.Integrate(dt) // updateObjects, ballistic.go:27-44: the backboard is never integrated
	cube.CalculateDerivedData()
	for _, b := range bullets {
		b.Body.Integrate(dt)
		b.CalculateDerivedData()
	}
	found := false // generateContacts, ballistic.go:47-97
	var f bool
	var contacts []*cubez.Contact
	ground := cubez.NewCollisionPlane(m.Vector3{0.0, 1.0, 0.0}, 0.0)
	f, contacts = cubez.CheckForCollisions(cube, ground, nil)
	if f {
		found = true
	}
	f, contacts = cubez.CheckForCollisions(cube, backboard, contacts)
	if f {
		found = true
	}
	for _, b := range bullets {
		f, contacts = cubez.CheckForCollisions(b, ground, contacts)
		if f {
			found = true
		}
		f, contacts = cubez.CheckForCollisions(cube, b, contacts)
		if f {
			found = true
		}
		f, contacts = cubez.CheckForCollisions(backboard, b, contacts)
		if f {
			found = true
		}
		for _, b2 := range bullets {
			if b2 == b {
				continue
			}
			f, contacts = cubez.CheckForCollisions(b2, b, contacts)
			if f {
				found = true
			}
		}
	}
	h := uint64(0xcbf29ce484222325)
	for _, c := range contacts {
		for k := 0; k < 2; k++ {
			v := uint64(0xFFFFFFFF)
			if c.Bodies[k] != nil {
				v = uint64(index[c.Bodies[k]])
			}
			h = (h ^ v) * 0x100000001b3
		}
		gen.vec3(&c.ContactPoint)
		gen.vec3(&c.ContactNormal)
		gen.real(c.Penetration)
	}
	*pairs += h
	if found {
		cubez.ResolveContacts(len(contacts)*8, contacts, dt) // ballistic.go:102-104
	}
	return len(contacts)
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
	steps := argInt(1, 600)
	nBullets := argInt(2, 64)
	first := argInt(3, 60)
	every := argInt(4, 8)
	setup()
	dt := m.Real(1.0 / 60.0)
	fmt.Printf("cubez-dump v2 scene=ballistic worlds=1 bodies=%d steps=%d\n", 2+nBullets, steps)
	// bodies that do not exist yet are dumped as the state they will be created with (what the world handle holds)
	var pending []*cubez.CollisionSphere
	saved := bullets
	for k := 0; k < nBullets; k++ {
		fire()
	}
	pending = bullets
	bullets = saved
	start := time.Now()
	for s := 0; s < steps; s++ {
		if len(bullets) < nBullets && s == first+every*len(bullets) {
			bullets = append(bullets, pending[len(bullets)])
		}
		var pairs uint64
		var gen hasher
		var state hasher
		total := step(dt, &pairs, &gen)
		state.body(cube.Body)
		state.body(backboard.Body)
		for _, b := range pending {
			state.body(b.Body)
		}
		fmt.Printf("step %d contacts %d pairhash %016x genhash %016x statehash %016x\n", s, total, pairs, gen.h, state.h)
	}
	wall := time.Since(start).Seconds()
	printBody(0, cube.Body)
	printBody(1, backboard.Body)
	for k, b := range pending {
		if k < 62 {
			printBody(2+k, b.Body)
		}
	}
	fmt.Fprintf(os.Stderr, "wall %.6f s for %d steps\n", wall, steps)
}
