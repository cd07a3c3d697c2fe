package cubez

/*
#include <stdlib.h>
#include "cubezcuda.h"
*/
import "C"

import (
	"unsafe"

	m "github.com/tbogdala/cubez/math"
)

// ---------------------------------------------------------------------------------------------
// Run — batched worlds sharded over several GPUs inside this one process (new API, cz_run_*): shard k of n owns worlds
// [k*W/n, (k+1)*W/n); stepping exchanges nothing between devices; Finish reduces checksum, energy and counters with
// one grouped ncclAllReduce (libnccl.so.2 is loaded by the library at NewRun).
// ---------------------------------------------------------------------------------------------
type Run struct {
	h                                         *C.cz_run
	NWorlds, BodiesPerWorld, ContactsPerWorld int
}

// RunTotals is cz_run_totals.
type RunTotals struct {
	Checksum                                           uint64
	Energy                                             float64
	WorldSteps, Contacts, PosIterations, VelIterations int64
	MaxDeviceMs                                        float32
	NShards                                            int
	UsedNCCL                                           bool
}

func runCheck(r *C.cz_run, rc C.int) {
	if rc != 0 {
		panic("cubez: " + C.GoString(C.cz_run_last_error(r)))
	}
}

// NewRun — cz_run_create: one shard per entry of devices; nWorlds is the TOTAL over all shards.
func NewRun(devices []int32, nWorlds, bodiesPerWorld, contactsPerWorld, schedule, flags int) *Run {
	n := len(devices)
	dp := cbuf(n, 4)
	defer C.free(dp)
	copy(unsafe.Slice((*int32)(dp), n), devices)
	d := C.cz_world_desc{n_worlds: C.int32_t(nWorlds), bodies_per_world: C.int32_t(bodiesPerWorld),
		contacts_per_world: C.int32_t(contactsPerWorld), schedule: C.int32_t(schedule), flags: C.int32_t(flags)}
	r := &Run{NWorlds: nWorlds, BodiesPerWorld: bodiesPerWorld, ContactsPerWorld: contactsPerWorld}
	runCheck(nil, C.cz_run_create(C.int32_t(n), (*C.int32_t)(dp), &d, &r.h))
	return r
}

// Close — cz_run_destroy.
func (r *Run) Close() {
	if r.h != nil {
		C.cz_run_destroy(r.h)
		r.h = nil
	}
}

// Shard — cz_run_shard: shard k's World (owned by the run: do not Close it) and the world range it holds.
func (r *Run) Shard(k int) (w *World, firstWorld, nWorlds int) {
	var h *C.cz_world
	var f, n C.int32_t
	runCheck(r.h, C.cz_run_shard(r.h, C.int32_t(k), &h, &f, &n))
	return &World{h: h, NWorlds: int(n), BodiesPerWorld: r.BodiesPerWorld, ContactsPerWorld: r.ContactsPerWorld}, int(f), int(n)
}

// UploadBodies / UploadColliders / UploadPlanes / SetEpisodes take whole-batch arrays and slice them per shard.
func (r *Run) UploadBodies(all *HostBodies, derive bool) {
	runCheck(r.h, C.cz_run_upload_bodies(r.h, &all.c, C.int32_t(b2u(derive))))
}
func (r *Run) UploadColliders(all *HostColliders, derive bool) {
	runCheck(r.h, C.cz_run_upload_colliders(r.h, &all.c, C.int32_t(b2u(derive))))
}
func (r *Run) UploadPlanes(planes []*CollisionPlane) {
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
	runCheck(r.h, C.cz_run_upload_planes(r.h, &p))
}
func (r *Run) SetEpisodes(length int, phase0 []int32) {
	var p *C.int32_t
	if phase0 != nil {
		q := cbuf(len(phase0), 4)
		defer C.free(q)
		copy(unsafe.Slice((*int32)(q), len(phase0)), phase0)
		p = (*C.int32_t)(q)
	}
	runCheck(r.h, C.cz_run_set_episodes(r.h, C.int32_t(length), p))
}

// Step — cz_run_step: n frames on every shard, enqueued on all devices before any is waited for.
func (r *Run) Step(dt m.Real, n int) { runCheck(r.h, C.cz_run_step(r.h, C.cz_real(dt), C.int32_t(n))) }

// Finish — cz_run_finish: wait for every shard, reduce, report; resets the counters and the timer.
func (r *Run) Finish() RunTotals {
	var t C.cz_run_totals
	runCheck(r.h, C.cz_run_finish(r.h, &t))
	return RunTotals{uint64(t.checksum), float64(t.energy), int64(t.world_steps), int64(t.contacts), int64(t.pos_iterations),
		int64(t.vel_iterations), float32(t.max_device_ms), int(t.n_shards), t.used_nccl != 0}
}
