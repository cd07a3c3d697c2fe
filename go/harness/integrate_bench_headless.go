// integrate_bench_headless: the UNMODIFIED reference (github.com/tbogdala/cubez) on BASELINE config 5 — n free rigid
// bodies, RigidBody.Integrate (+ CalculateDerivedData inside it, rigidbody.go:213-299) and nothing else: the CPU
// baseline of the HBM-roofline microbench, and a per-frame state hash for parity.
//
//	go run integrate_bench_headless.go <steps> [bodies [seed]] > dump.txt ; python tools/compare_go_dump.py dump.txt
//
// Note: this scene draws a different damping per body, so math.Pow (rigidbody.go:233-234) is evaluated with 2n
// different bases — the one place where Go's Pow and C's pow() may differ by an ulp (the library takes the Pow factors
// as host inputs for that reason: cz_world_set_pow / cz_integrate).
// Also translated by oracle/go2cpp.py into oracle/_ref/integrate_bench_headless: keep it inside that tool's Go subset.
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

// SURVEY section 8d, cfg5: splitmix64(seed*0x100000001B3 + i*0x9E3779B97F4A7C15), 21 draws per body
func build(n int, seed uint64) []*cubez.RigidBody {
	var bodies []*cubez.RigidBody
	for i := 0; i < n; i++ {
		state := seed*0x100000001B3 + uint64(i)*0x9E3779B97F4A7C15
		var u [21]float64
		for k := 0; k < 21; k++ {
			u[k] = draw(&state)
		}
		b := cubez.NewRigidBody()
		b.Position = m.Vector3{m.Real(uniform(u[0], -100.0, 100.0)), m.Real(uniform(u[1], -100.0, 100.0)), m.Real(uniform(u[2], -100.0, 100.0))}
		q0 := uniform(u[3], -1.0, 1.0)
		q1 := uniform(u[4], -1.0, 1.0)
		q2 := uniform(u[5], -1.0, 1.0)
		q3 := uniform(u[6], -1.0, 1.0)
		if math.Sqrt(q0*q0+q1*q1+q2*q2+q3*q3) < 0.1 {
			q0 = 1.0
			q1 = 0.0
			q2 = 0.0
			q3 = 0.0
		}
		q := m.Quat{m.Real(q0), m.Real(q1), m.Real(q2), m.Real(q3)}
		ln := m.RealSqrt(q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3])
		b.Orientation = m.Quat{q[0] / ln, q[1] / ln, q[2] / ln, q[3] / ln}
		b.Velocity = m.Vector3{m.Real(uniform(u[7], -5.0, 5.0)), m.Real(uniform(u[8], -5.0, 5.0)), m.Real(uniform(u[9], -5.0, 5.0))}
		b.Rotation = m.Vector3{m.Real(uniform(u[10], -3.0, 3.0)), m.Real(uniform(u[11], -3.0, 3.0)), m.Real(uniform(u[12], -3.0, 3.0))}
		b.LinearDamping = m.Real(uniform(u[13], 0.90, 0.99))
		b.AngularDamping = m.Real(uniform(u[14], 0.90, 0.99))
		var t m.Matrix3
		t[0] = m.Real(uniform(u[15], 0.5, 2.0))
		t[4] = m.Real(uniform(u[16], 0.5, 2.0))
		t[8] = m.Real(uniform(u[17], 0.5, 2.0))
		t[1] = m.Real(uniform(u[18], -0.1, 0.1))
		t[3] = t[1]
		t[2] = m.Real(uniform(u[19], -0.1, 0.1))
		t[6] = t[2]
		t[5] = m.Real(uniform(u[20], -0.1, 0.1))
		t[7] = t[5]
		b.InverseInertiaTensor = t
		b.SetMass(1.0)
		b.CalculateDerivedData()
		bodies = append(bodies, b)
	}
	return bodies
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
	steps := argInt(1, 10)
	n := argInt(2, 65536)
	seed := argInt(3, 5)
	bodies := build(n, uint64(seed))
	dt := m.Real(1.0 / 60.0)
	fmt.Printf("cubez-dump v2 scene=free_bodies worlds=1 bodies=%d steps=%d\n", n, steps)
	start := time.Now()
	var inLoop float64
	for s := 0; s < steps; s++ {
		t0 := time.Now()
		for _, b := range bodies {
			b.Integrate(dt)
		}
		inLoop += time.Since(t0).Seconds()
		var state hasher
		for _, b := range bodies {
			state.body(b)
		}
		fmt.Printf("step %d contacts 0 pairhash cbf29ce484222325 genhash 0000000000000000 statehash %016x\n", s, state.h)
	}
	wall := time.Since(start).Seconds()
	for i, b := range bodies {
		if i < 64 {
			printBody(i, b)
		}
	}
	fmt.Fprintf(os.Stderr, "wall %.6f s (%.6f s inside Integrate) for %d steps of %d bodies: %.0f body-steps/s\n", wall, inLoop, steps, n, float64(n)*float64(steps)/inLoop)
}
