// cubedrop_headless: runs the UNMODIFIED reference (github.com/tbogdala/cubez) on the cfg1 scene
// (examples/cubedrop.go:146-177 fire() twice, ground plane, dt = 1/60) without GL and prints one
// line per frame in the same format as tests/golden (contact count, pair-sequence FNV hash) plus the
// final state, so the oracle <-> Go gap can be closed by anyone with a Go toolchain:
//
//	go run cubedrop_headless.go 200 > go_cubedrop.txt ; python tools/compare_go_dump.py go_cubedrop.txt
//
// SOURCE ONLY here (no Go in the build image).  Straight-line calls of the public API only.
package main

import (
	"fmt"
	gomath "math"
	"os"
	"strconv"
	"time"

	"github.com/tbogdala/cubez"
	m "github.com/tbogdala/cubez/math"
)

var cubes []*cubez.CollisionCube

func fire() { // examples/cubedrop.go:146-177 without the GL node
	var offset float32
	if len(cubes) > 0 && (len(cubes)/4)%2 >= 1 {
		offset = 0.75
	}
	for i := 0; i < 4; i++ {
		c := cubez.NewCollisionCube(nil, m.Vector3{0.5, 0.5, 0.5})
		c.Body.Position = m.Vector3{m.Real(i*2.0-4/2) - 0.5 + m.Real(offset), 10.0, 0.0}
		c.Body.SetMass(8.0)
		c.Body.CanSleep = true
		var inertia m.Matrix3
		inertia.SetBlockInertiaTensor(&c.HalfSize, 8.0)
		c.Body.SetInertiaTensor(&inertia)
		c.Body.CalculateDerivedData()
		c.CalculateDerivedData()
		cubes = append(cubes, c)
	}
}

func main() {
	steps := 600
	if len(os.Args) > 1 {
		steps, _ = strconv.Atoi(os.Args[1])
	}
	ground := cubez.NewCollisionPlane(m.Vector3{0.0, 1.0, 0.0}, 0.0)
	fire()
	fire()
	index := map[*cubez.RigidBody]int{}
	for i, c := range cubes {
		index[c.Body] = i
	}
	delta := 1.0 / 60.0
	start := time.Now()
	for s := 0; s < steps; s++ {
		for _, c := range cubes { // updateObjects, cubedrop.go:29-39
			c.Body.Integrate(m.Real(delta))
			c.CalculateDerivedData()
		}
		var contacts []*cubez.Contact // generateContacts, cubedrop.go:42-67
		found := false
		for _, c := range cubes {
			var f bool
			f, contacts = c.CheckAgainstHalfSpace(ground, contacts)
			found = found || f
			for _, o := range cubes {
				if o == c {
					continue
				}
				f, contacts = cubez.CheckForCollisions(c, o, contacts)
				found = found || f
			}
		}
		h := uint64(0xcbf29ce484222325) // FNV-1a over the (body0, body1) sequence, as tests/hostemu_lib.pair_hash
		for _, c := range contacts {
			for k := 0; k < 2; k++ {
				v := uint64(0xFFFFFFFF)
				if c.Bodies[k] != nil {
					v = uint64(index[c.Bodies[k]])
				}
				h = (h ^ v) * 0x100000001b3
			}
		}
		fmt.Printf("step %d contacts %d pairhash %016x\n", s, len(contacts), h)
		if found {
			cubez.ResolveContacts(len(contacts)*8, contacts, m.Real(delta))
		}
	}
	fmt.Fprintf(os.Stderr, "wall %.6f s for %d steps\n", time.Since(start).Seconds(), steps)
	for i, c := range cubes {
		b := c.Body
		fmt.Printf("body %d pos %x %x %x ori %x %x %x %x vel %x %x %x rot %x %x %x awake %v\n", i,
			mathBits(b.Position[0]), mathBits(b.Position[1]), mathBits(b.Position[2]),
			mathBits(b.Orientation[0]), mathBits(b.Orientation[1]), mathBits(b.Orientation[2]), mathBits(b.Orientation[3]),
			mathBits(b.Velocity[0]), mathBits(b.Velocity[1]), mathBits(b.Velocity[2]),
			mathBits(b.Rotation[0]), mathBits(b.Rotation[1]), mathBits(b.Rotation[2]), b.IsAwake)
	}
}

func mathBits(r m.Real) uint64 { return gomath.Float64bits(float64(r)) }
