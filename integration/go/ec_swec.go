// NOTE: this file has never been compiled — the build image has no Go toolchain.  It is reviewed source, written
// against include/swec.h; the same C ABI is exercised from C (tests/c/cgo_shaped_harness.c) and Python on the GPU.
//go:build swec && cgo

// weed/storage/erasure_coding/ec_swec.go
package erasure_coding

/*
#cgo CFLAGS:  -I${SRCDIR}/../../../third_party/swec/include
#cgo LDFLAGS: -L${SRCDIR}/../../../third_party/swec/lib -lswec -lstdc++ -ldl -lpthread
#include <stdlib.h>
#include "swec.h"
*/
import "C"

import (
	"fmt"
	"io"
	"runtime"
	"sync/atomic"
	"unsafe"

	"github.com/klauspost/reedsolomon"
)

// One volume server process drives every GPU of the box: each encoder / file-level call takes the next
// GPU round-robin (volume v → GPU v mod N — independent volumes need no collective), the same way the
// shell already runs up to 10 volumes concurrently (weed/shell/common.go:11).  The round-robin walks the
// library's socket-interleaved order (0,4,1,5,… on a two-socket box): n concurrent volumes then use both
// sockets' memory controllers — one socket cannot feed four GPUs at full PCIe rate.  Set -ec.gpu=N to pin one.
var (
	swecPinned  = -1 // -ec.gpu flag; -1 = round-robin over all devices
	swecNext    uint32
	swecDevices = func() []C.int {
		var order [64]C.int
		var n C.int
		if C.swec_device_spread_order(&order[0], 64, &n) != C.SWEC_OK || n < 1 {
			return []C.int{0} // calls will fail with SWEC_ERR_NO_DEVICE and surface as Go errors
		}
		return append([]C.int(nil), order[:int(n)]...)
	}()
)

func swecPickDevice() C.int {
	if swecPinned >= 0 {
		return C.int(swecPinned)
	}
	return swecDevices[int(atomic.AddUint32(&swecNext, 1))%len(swecDevices)]
}

type swecEncoder struct {
	h          *C.swec_encoder
	data, par  int
}

// swecCall runs one C entry point and, on failure, fetches its detail string.  swec_last_error() is
// thread-local and the Go scheduler may move a goroutine to another OS thread BETWEEN two cgo calls, so the
// failing call and the read of its detail are bracketed by LockOSThread.
func swecCall(f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	return swecErr(f())
}

// swecErr must run on the OS thread that made the failing call (see swecCall).
func swecErr(rc C.int) error {
	if rc == C.SWEC_OK {
		return nil
	}
	msg := C.GoString(C.swec_last_error())
	switch rc {
	case C.SWEC_ERR_TOO_FEW_SHARDS:
		return reedsolomon.ErrTooFewShards
	case C.SWEC_ERR_SHARD_SIZE:
		return reedsolomon.ErrShardSize
	}
	return fmt.Errorf("swec: %s: %s", C.GoString(C.swec_strerror(rc)), msg)
}

// newSwecEncoder replaces reedsolomon.New(ds, ps) (ec_context.go:35, store_ec.go:485).
func newSwecEncoder(dataShards, parityShards int) (reedsolomon.Encoder, error) {
	var h *C.swec_encoder
	if err := swecCall(func() C.int {
		return C.swec_encoder_new(C.int(dataShards), C.int(parityShards), swecPickDevice(), &h)
	}); err != nil {
		return nil, err
	}
	e := &swecEncoder{h: h, data: dataShards, par: parityShards}
	runtime.SetFinalizer(e, func(e *swecEncoder) { C.swec_encoder_free(e.h) })
	return e, nil
}

// pin builds the C pointer table. The slices' backing arrays are Go memory: cgo forbids storing Go
// pointers in C memory across calls, but passing a C array of Go pointers for the duration of one
// call is allowed with runtime.Pinner (Go ≥ 1.21).
func (e *swecEncoder) pin(shards [][]byte, p *runtime.Pinner) (**C.uint8_t, func()) {
	n := e.data + e.par
	tbl := (*[1 << 10]*C.uint8_t)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	for i := 0; i < n; i++ {
		if len(shards[i]) > 0 {
			p.Pin(&shards[i][0])
			tbl[i] = (*C.uint8_t)(unsafe.Pointer(&shards[i][0]))
		} else {
			tbl[i] = nil
		}
	}
	return (**C.uint8_t)(unsafe.Pointer(tbl)), func() { C.free(unsafe.Pointer(tbl)) }
}

func shardLen(shards [][]byte) (int, error) {
	n := 0
	for _, s := range shards {
		if len(s) == 0 {
			continue
		}
		if n == 0 {
			n = len(s)
		} else if len(s) != n {
			return 0, reedsolomon.ErrShardSize
		}
	}
	if n == 0 {
		return 0, reedsolomon.ErrShardNoData
	}
	return n, nil
}

// Encode: parity slices overwritten in place, data untouched (ec_encoder.go:265).
func (e *swecEncoder) Encode(shards [][]byte) error {
	if len(shards) != e.data+e.par {
		return reedsolomon.ErrTooFewShards
	}
	n, err := shardLen(shards)
	if err != nil {
		return err
	}
	for _, s := range shards {
		if len(s) != n {
			return reedsolomon.ErrShardSize
		}
	}
	var p runtime.Pinner
	defer p.Unpin()
	tbl, free := e.pin(shards, &p)
	defer free()
	return swecCall(func() C.int { return C.swec_encode(e.h, tbl, C.size_t(n)) })
}

func (e *swecEncoder) reconstruct(shards [][]byte, dataOnly bool) error {
	if len(shards) != e.data+e.par {
		return reedsolomon.ErrTooFewShards
	}
	n, err := shardLen(shards)
	if err != nil {
		return err
	}
	present := make([]C.uint8_t, len(shards))
	have := 0
	for i, s := range shards {
		if len(s) > 0 {
			present[i] = 1
			have++
		}
	}
	if have == len(shards) {
		return nil
	}
	if have < e.data {
		return reedsolomon.ErrTooFewShards
	}
	for i := range shards { // klauspost allocates nil/empty shards (re-using capacity when it can)
		if present[i] == 0 && (i < e.data || !dataOnly) {
			if cap(shards[i]) >= n {
				shards[i] = shards[i][:n]
			} else {
				shards[i] = make([]byte, n)
			}
		}
	}
	var p runtime.Pinner
	defer p.Unpin()
	tbl, free := e.pin(shards, &p)
	defer free()
	d := C.int(0)
	if dataOnly {
		d = 1
	}
	return swecCall(func() C.int { return C.swec_reconstruct(e.h, tbl, &present[0], C.size_t(n), d) })
}

func (e *swecEncoder) Reconstruct(shards [][]byte) error     { return e.reconstruct(shards, false) } // ec_encoder.go:360
func (e *swecEncoder) ReconstructData(shards [][]byte) error { return e.reconstruct(shards, true) }  // store_ec.go:551

func (e *swecEncoder) Verify(shards [][]byte) (bool, error) {
	n, err := shardLen(shards)
	if err != nil {
		return false, err
	}
	var p runtime.Pinner
	defer p.Unpin()
	tbl, free := e.pin(shards, &p)
	defer free()
	var ok C.int
	if err := swecCall(func() C.int { return C.swec_verify(e.h, tbl, C.size_t(n), &ok) }); err != nil {
		return false, err
	}
	return ok != 0, nil
}

// The remaining reedsolomon.Encoder methods are not used by SeaweedFS on this path
// (grep: only Encode, Reconstruct, ReconstructData are called); they return ErrNotSupported.
func (e *swecEncoder) EncodeIdx([]byte, int, [][]byte) error          { return reedsolomon.ErrNotSupported }
func (e *swecEncoder) ReconstructSome([][]byte, []bool) error         { return reedsolomon.ErrNotSupported }
func (e *swecEncoder) Update([][]byte, [][]byte) error                { return reedsolomon.ErrNotSupported }
func (e *swecEncoder) Split([]byte) ([][]byte, error)                 { return nil, reedsolomon.ErrNotSupported }
func (e *swecEncoder) Join(w io.Writer, s [][]byte, n int) error      { return reedsolomon.ErrNotSupported }

// ---- file-level entry points (preferred: one cgo crossing per volume instead of 12,288) ----------

// generateEcFilesSwec replaces the body of generateEcFiles (ec_encoder.go:110-128).
func generateEcFilesSwec(baseFileName string, bufferSize int, largeBlockSize, smallBlockSize int64, ctx *ECContext) error {
	cs := C.CString(baseFileName)
	defer C.free(unsafe.Pointer(cs))
	if err := swecCall(func() C.int {
		return C.swec_generate_ec_files(cs, C.int64_t(bufferSize), C.int64_t(largeBlockSize), C.int64_t(smallBlockSize),
			C.int(ctx.DataShards), C.int(ctx.ParityShards), swecPickDevice())
	}); err != nil {
		return fmt.Errorf("encodeDatFile: %w", err)
	}
	return nil
}

// generateMissingEcFilesSwec replaces generateMissingEcFiles (ec_encoder.go:146-200).
func generateMissingEcFilesSwec(baseFileName string, ctx *ECContext, additionalDirs []string) ([]uint32, error) {
	cs := C.CString(baseFileName)
	defer C.free(unsafe.Pointer(cs))
	dirs := make([]*C.char, len(additionalDirs)+1)
	for i, d := range additionalDirs {
		dirs[i] = C.CString(d)
		defer C.free(unsafe.Pointer(dirs[i]))
	}
	var rebuilt [C.SWEC_MAX_SHARDS]C.uint32_t
	var n C.int
	if err := swecCall(func() C.int {
		return C.swec_rebuild_ec_files(cs, (**C.char)(unsafe.Pointer(&dirs[0])), C.int(len(additionalDirs)),
			C.int(ctx.DataShards), C.int(ctx.ParityShards), swecPickDevice(), &rebuilt[0], &n)
	}); err != nil {
		return nil, fmt.Errorf("rebuildEcFiles: %w", err)
	}
	ids := make([]uint32, int(n))
	for i := range ids {
		ids[i] = uint32(rebuilt[i])
	}
	return ids, nil
}

// ---- volume-level entry points: the file work of the three EC gRPC handlers, one cgo crossing each --------

// ecShardsGenerateSwec replaces the file work of VolumeEcShardsGenerate (weed/server/volume_grpc_erasure_coding.go:43-146):
// EC ratio from an existing .vif, .ecx before the shards, .dat size snapshot, shards on the GPU, .vif, cleanup on error.
// The handler keeps the volume lookup, the collection check and the maintenance-mode check.
func ecShardsGenerateSwec(dataBaseFileName, indexBaseFileName string, needleVersion uint32, expireAtSec uint64) error {
	cd, ci := C.CString(dataBaseFileName), C.CString(indexBaseFileName)
	defer C.free(unsafe.Pointer(cd))
	defer C.free(unsafe.Pointer(ci))
	return swecCall(func() C.int {
		return C.swec_ec_shards_generate(cd, ci, C.uint32_t(needleVersion), C.uint64_t(expireAtSec), swecPickDevice())
	})
}

// ecShardsRebuildSwec replaces RebuildEcFiles + RebuildEcxFile in VolumeEcShardsRebuild (:149-225).
func ecShardsRebuildSwec(dataBaseFileName, indexBaseFileName string, additionalDirs []string) ([]uint32, error) {
	cd, ci := C.CString(dataBaseFileName), C.CString(indexBaseFileName)
	defer C.free(unsafe.Pointer(cd))
	defer C.free(unsafe.Pointer(ci))
	dirs := make([]*C.char, len(additionalDirs)+1)
	for i, d := range additionalDirs {
		dirs[i] = C.CString(d)
		defer C.free(unsafe.Pointer(dirs[i]))
	}
	var rebuilt [C.SWEC_MAX_SHARDS]C.uint32_t
	var n C.int
	if err := swecCall(func() C.int {
		return C.swec_ec_shards_rebuild(cd, ci, (**C.char)(unsafe.Pointer(&dirs[0])), C.int(len(additionalDirs)),
			swecPickDevice(), &rebuilt[0], &n)
	}); err != nil {
		return nil, err
	}
	ids := make([]uint32, int(n))
	for i := range ids {
		ids[i] = uint32(rebuilt[i])
	}
	return ids, nil
}

// ecShardsToVolumeSwec replaces the file work of VolumeEcShardsToVolume (:578-668); errNoLiveEntries maps to the
// handler's FailedPrecondition with EcNoLiveEntriesSubstring.  Compaction stays in Go.
var errNoLiveEntries = fmt.Errorf("ec volume %s", EcNoLiveEntriesSubstring)

func ecShardsToVolumeSwec(dataBaseFileName, indexBaseFileName string, additionalDirs []string) (datFileSize int64, err error) {
	cd, ci := C.CString(dataBaseFileName), C.CString(indexBaseFileName)
	defer C.free(unsafe.Pointer(cd))
	defer C.free(unsafe.Pointer(ci))
	dirs := make([]*C.char, len(additionalDirs)+1)
	for i, d := range additionalDirs {
		dirs[i] = C.CString(d)
		defer C.free(unsafe.Pointer(dirs[i]))
	}
	var size C.int64_t
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	rc := C.swec_ec_shards_to_volume(cd, ci, (**C.char)(unsafe.Pointer(&dirs[0])), C.int(len(additionalDirs)), &size)
	if rc == C.SWEC_ERR_NO_LIVE_NEEDLES {
		return 0, errNoLiveEntries
	}
	return int64(size), swecErr(rc)
}

// ---- the read path of a mounted EC volume whose shards are local files ----------------------------------------

// swecEcVolume is the twin of the long-lived EcVolume (ec_volume.go:36-160) for Store.ReadEcShardNeedle
// (weed/storage/store_ec.go:252-355): one handle per mounted volume, closed on unmount.
type swecEcVolume struct{ h *C.swec_ec_volume }

func openSwecEcVolume(dataBaseFileName, indexBaseFileName string, additionalDirs []string) (*swecEcVolume, error) {
	cd, ci := C.CString(dataBaseFileName), C.CString(indexBaseFileName)
	defer C.free(unsafe.Pointer(cd))
	defer C.free(unsafe.Pointer(ci))
	dirs := make([]*C.char, len(additionalDirs)+1)
	for i, d := range additionalDirs {
		dirs[i] = C.CString(d)
		defer C.free(unsafe.Pointer(dirs[i]))
	}
	var h *C.swec_ec_volume
	if err := swecCall(func() C.int {
		return C.swec_ec_volume_open(cd, ci, (**C.char)(unsafe.Pointer(&dirs[0])), C.int(len(additionalDirs)), swecPickDevice(), &h)
	}); err != nil {
		return nil, err
	}
	return &swecEcVolume{h: h}, nil
}

func (v *swecEcVolume) Close() { C.swec_ec_volume_close(v.h); v.h = nil }

// ReadNeedles returns the raw record bytes of every id (nil where the needle is unknown or deleted), reading present
// shards directly and rebuilding the intervals that lived on lost shards in ONE batched GPU call.  bufs are C memory
// (C.malloc), because the call writes through pointers stored in a C array.
func (v *swecEcVolume) ReadNeedles(ids []uint64, capacity int) ([][]byte, []error, error) {
	n := len(ids)
	if n == 0 {
		return nil, nil, nil
	}
	reads := (*[1 << 20]C.swec_needle_read)(C.calloc(C.size_t(n), C.size_t(unsafe.Sizeof(C.swec_needle_read{}))))[:n:n]
	defer C.free(unsafe.Pointer(&reads[0]))
	arena := C.malloc(C.size_t(n * capacity))
	defer C.free(arena)
	for i, id := range ids {
		reads[i].needle_id = C.uint64_t(id)
		reads[i].buf = (*C.uint8_t)(unsafe.Add(arena, i*capacity))
		reads[i].capacity = C.size_t(capacity)
	}
	if err := swecCall(func() C.int { return C.swec_ec_volume_read_needles(v.h, &reads[0], C.int(n)) }); err != nil {
		return nil, nil, err
	}
	out, errs := make([][]byte, n), make([]error, n)
	for i := range reads {
		switch reads[i].status {
		case C.SWEC_OK:
			out[i] = C.GoBytes(unsafe.Pointer(reads[i].buf), C.int(reads[i].n_bytes))
		case C.SWEC_ERR_NOT_FOUND:
			errs[i] = NotFoundError
		case C.SWEC_ERR_DELETED:
			errs[i] = fmt.Errorf("already deleted") // storage.ErrorDeleted at the call site
		default:
			errs[i] = fmt.Errorf("swec: needle %x: %s", ids[i], C.GoString(C.swec_strerror(C.int(reads[i].status))))
		}
	}
	return out, errs, nil
}

// DeleteNeedleFromEcx: journal append (ec_volume_delete.go:28-93).
func (v *swecEcVolume) DeleteNeedleFromEcx(id uint64) error {
	return swecCall(func() C.int { return C.swec_ec_volume_delete_needle(v.h, C.uint64_t(id)) })
}

// ---- pinned batch buffers ---------------------------------------------------------------------------------
// runtime.Pinner only stops the Go GC from moving a slice; to CUDA such memory is PAGEABLE, so every Encode /
// Reconstruct on it bounces through the library's pinned ring (a memcpy per shard each way).  The batch buffers of
// the two file loops are allocated once per volume and reused for every batch, so they are the place to hand the
// library memory it can DMA — or read from the kernel — directly:
//
//   encodeDatFile      ec_encoder.go:289-292   buffers[i] = make([]byte, bufferSize)   ×TotalShards
//   rebuildEcFiles     ec_encoder.go:330-335   buffers[i] = make([]byte, ErasureCodingSmallBlockSize) ×TotalShards
//   recoverOneRemoteEcShardInterval  store_ec.go:493-500   bufs[i] = make([]byte, len(buf))  (per degraded read)
//
// become   buffers, release := swecAllocShardBuffers(ctx.Total(), bufferSize); defer release()
//
// The n slices are cut from ONE pinned, GPU-mapped allocation on the GPU's NUMA node at a constant pitch: the library
// recognises that shape and moves all k inputs (and all m outputs) with one strided DMA each way, or — for calls up to
// 2 MiB per shard — runs the kernel directly on the host memory over PCIe (include/swec.h "host_zero_copy").
// The memory is C memory: no Pinner, no cgo pointer-passing rules, and the slices must not outlive release().
func swecAllocShardBuffers(n int, shardLen int) (bufs [][]byte, release func()) {
	pitch := (shardLen + 4095) &^ 4095
	dev := swecPickDevice()
	base := C.swec_alloc_pinned_for_device(dev, C.size_t(n*pitch))
	if base == nil { // no GPU / out of pinned memory: plain Go memory still works (pageable path)
		bufs = make([][]byte, n)
		for i := range bufs {
			bufs[i] = make([]byte, shardLen)
		}
		return bufs, func() {}
	}
	bufs = make([][]byte, n)
	for i := range bufs {
		bufs[i] = unsafe.Slice((*byte)(unsafe.Add(base, i*pitch)), shardLen)
	}
	return bufs, func() { C.swec_free_pinned(base) }
}
