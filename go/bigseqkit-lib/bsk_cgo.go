// Package main -- drop-in operators for the BigSeqKit executor plugin (bigseqkit.so) that run the per-record
// hot path on a B200 through libbsk.so (include/bsk.h).  NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Go
// toolchain, IgnisHPC absent): it is the binding a BigSeqKit maintainer adds next to bigseqkit-lib/*.go.
//
// The plugin keeps the reference's exported factories with their element types (bigseqkit-lib/seq.go:17-19
// NewSeqTransform, stats.go:16 NewStats, stats.go:119 NewStatsReduce, rmdup.go:23 NewRmDupPrepare, translate.go:21
// NewTranslate, locate.go:19 NewLocate, grep.go:24 NewGrep, subseq.go:22 NewSubseqTransform, fq2fa.go:15 NewFq2Fa),
// the same embedded base.I... types and the same "opts" JSON variable (bigseqkit/helper.go:47-66), so IgnisHPC and
// the UNCHANGED drivers (bigseqkit/, bigseqkit-py/, bigseqkit-cli/) see an identical plugin:
//   stats   bigseqkit/stats.go:86-94   MapPartitions(Stats) -> Reduce(StatsReduce): both defined here;
//   rmdup   bigseqkit/rmdup.go:92-108  MapPartitions(RmDupPrepare) -> GroupByKey -> Flatmap(RmDupCheck):
//           RmDupPrepare (the pass that parses, hashes and formats every record) is defined here and returns
//           IPair[int64, string] like the reference; RmDupCheck stays the reference's own Go code
//           (bigseqkit-lib/rmdup.go:92-275 is kept in the plugin: per-group host logic on groups of 1-2 strings).
// NewRmDupSharded at the end is the optional faster rmdup (one exchange of 16-byte fingerprints over NCCL instead of
// shuffling every record through GroupByKey); it needs the three-line driver patch shown in INTEGRATION.md.
//
// Build (inside ignishpc/go-compiler, with libbsk.so and bsk.h installed):
//   CGO_CFLAGS="-I/opt/bsk/include" CGO_LDFLAGS="-L/opt/bsk/lib -lbsk" go build -buildmode=plugin -trimpath -o bigseqkit.so
package main

/*
#include <stdlib.h>
#include "bsk.h"
*/
import "C"

import (
	"fmt"
	"os"
	"path"
	"strconv"
	"unsafe"

	"bigseqkit"

	"ignis/executor/api"
	"ignis/executor/api/base"
	"ignis/executor/api/ipair"
	"ignis/executor/api/iterator"
)

// bskOp is one operator instance between Before() and After(): one bsk_ctx per executor thread, because the
// reference calls Call() concurrently, once per partition, from ctx.Threads() threads (bigseqkit-lib/helper.go:413-416).
type bskOp struct {
	name string
	ctxs []*C.bsk_ctx
}

func (o *bskOp) before(context api.IContext, name string) error {
	o.name = name
	opts := C.CString(context.Vars()["opts"].(string)) // the reference's own JSON, defaults filled by the driver
	defer C.free(unsafe.Pointer(opts))
	cname := C.CString(name)
	defer C.free(unsafe.Pointer(cname))
	nGPU := int(C.bsk_device_count())
	if nGPU < 1 {
		return fmt.Errorf("bigseqkit-b200: no CUDA device visible to executor %d", context.ExecutorId())
	}
	dev := C.int(context.ExecutorId() % nGPU)
	o.ctxs = make([]*C.bsk_ctx, context.Threads())
	for i := range o.ctxs {
		if rc := C.bsk_create(cname, opts, dev, &o.ctxs[i]); rc != C.BSK_OK {
			return fmt.Errorf("%s", C.GoString(C.bsk_create_error())) // same text as the reference's Before()
		}
	}
	return nil
}

func (o *bskOp) after() error {
	for _, c := range o.ctxs {
		C.bsk_destroy(c)
	}
	o.ctxs = nil
	return nil
}

// call runs one partition: the iterator's record strings are packed into one pinned arena ('\n'-separated, exactly
// the bytes PlainFile + ReadFixer produced), one bsk_run_buffer call does the work, and the returned arena is
// sliced into Go strings by the element offsets.  One cgo call per partition, not per record.
func (o *bskOp) call(pid int64, it iterator.IReadIterator[string], context api.IContext) ([]string, error) {
	ctx := o.ctxs[context.ThreadId()]
	buf := make([]byte, 0, 64<<20)
	for it.HasNext() {
		e, err := it.Next()
		if err != nil {
			return nil, err
		}
		buf = append(buf, e...)
		buf = append(buf, '\n')
	}
	var out C.bsk_out
	var p *C.uint8_t
	if len(buf) > 0 {
		p = (*C.uint8_t)(unsafe.Pointer(&buf[0]))
	}
	if rc := C.bsk_run_buffer(ctx, p, C.size_t(len(buf)), C.int64_t(pid), &out); rc != C.BSK_OK {
		return nil, fmt.Errorf("%s", C.GoString(C.bsk_last_error(ctx)))
	}
	n := int(out.n_elem)
	res := make([]string, n)
	if n > 0 {
		data := unsafe.Slice((*byte)(unsafe.Pointer(out.data)), int(out.n))
		off := unsafe.Slice((*uint64)(unsafe.Pointer(out.elem_off)), n+1)
		for i := 0; i < n; i++ {
			res[i] = string(data[off[i] : off[i+1]-1]) // element without the '\n' FileStore adds back
		}
	}
	return res, nil
}

// callSharded: the partition goes to HBM once (bsk_run_buffer on a Prepare-free path is not enough here: the exchange
// needs the shard resident), bsk_rmdup_sharded does hash -> all-gather -> resolve, the survivors come back through
// bsk_memcpy_d2h.  Device staging is done by libbsk (bsk_stage_device) so that no CUDA call appears in Go.
func (o *bskOp) callSharded(it iterator.IReadIterator[string], context api.IContext) ([]string, error) {
	ctx := o.ctxs[0]
	buf := make([]byte, 0, 64<<20)
	for it.HasNext() {
		e, err := it.Next()
		if err != nil {
			return nil, err
		}
		buf = append(buf, e...)
		buf = append(buf, '\n')
	}
	var dptr unsafe.Pointer
	var p *C.uint8_t
	if len(buf) > 0 {
		p = (*C.uint8_t)(unsafe.Pointer(&buf[0]))
	}
	if rc := C.bsk_stage_device(ctx, p, C.size_t(len(buf)), &dptr); rc != C.BSK_OK {
		return nil, fmt.Errorf("%s", C.GoString(C.bsk_last_error(ctx)))
	}
	var out C.bsk_out
	if rc := C.bsk_rmdup_sharded(ctx, dptr, C.size_t(len(buf)), &out); rc != C.BSK_OK {
		return nil, fmt.Errorf("%s", C.GoString(C.bsk_last_error(ctx)))
	}
	data := make([]byte, int(out.n))
	off := make([]uint64, int(out.n_elem)+1)
	if out.n > 0 {
		C.bsk_memcpy_d2h(ctx, unsafe.Pointer(&data[0]), unsafe.Pointer(out.data), out.n)
		C.bsk_memcpy_d2h(ctx, unsafe.Pointer(&off[0]), unsafe.Pointer(out.elem_off), C.size_t(8*len(off)))
	}
	res := make([]string, int(out.n_elem))
	for i := range res {
		res[i] = string(data[off[i] : off[i+1]-1])
	}
	return res, nil
}

// ---- SeqTransform: bigseqkit-lib/seq.go:17-269
func NewSeqTransform() any { return &SeqTransform{} }

type SeqTransform struct {
	base.IMapPartitions[string, string]
	op bskOp
}

func (t *SeqTransform) Before(context api.IContext) error { return t.op.before(context, "SeqTransform") }
func (t *SeqTransform) After(context api.IContext) error { return t.op.after() } // frees the executor's ctxs
func (t *SeqTransform) Call(it iterator.IReadIterator[string], context api.IContext) ([]string, error) {
	return t.op.call(0, it, context)
}

// ---- SubseqTransform (region mode): bigseqkit-lib/subseq.go:22-526
func NewSubseqTransform() any { return &SubseqTransform{} }

type SubseqTransform struct {
	base.IMapPartitions[string, string]
	op bskOp
}

func (t *SubseqTransform) Before(context api.IContext) error { return t.op.before(context, "SubseqTransform") }
func (t *SubseqTransform) After(context api.IContext) error { return t.op.after() } // frees the executor's ctxs
func (t *SubseqTransform) Call(it iterator.IReadIterator[string], context api.IContext) ([]string, error) {
	return t.op.call(0, it, context)
}

// ---- Fq2Fa: bigseqkit-lib/fq2fa.go:15-61 (first of the SURVEY section 8f "next" operators)
func NewFq2Fa() any { return &Fq2Fa{} }

type Fq2Fa struct {
	base.IMapPartitions[string, string]
	op bskOp
}

func (t *Fq2Fa) Before(context api.IContext) error { return t.op.before(context, "Fq2Fa") }
func (t *Fq2Fa) After(context api.IContext) error { return t.op.after() } // frees the executor's ctxs
func (t *Fq2Fa) Call(it iterator.IReadIterator[string], context api.IContext) ([]string, error) {
	return t.op.call(0, it, context)
}

// ---- Translate: bigseqkit-lib/translate.go:21-145
func NewTranslate() any { return &Translate{} }

type Translate struct {
	base.IMapPartitions[string, string]
	op bskOp
}

func (t *Translate) Before(context api.IContext) error { return t.op.before(context, "Translate") }
func (t *Translate) After(context api.IContext) error { return t.op.after() } // frees the executor's ctxs
func (t *Translate) Call(it iterator.IReadIterator[string], context api.IContext) ([]string, error) {
	return t.op.call(0, it, context)
}

// ---- Locate / Grep: IMapPartitionsWithIndex (bigseqkit-lib/locate.go:19,195 ; grep.go:24,544)
func NewLocate() any { return &Locate{} }

type Locate struct {
	base.IMapPartitionsWithIndex[string, string]
	op bskOp
}

func (t *Locate) Before(context api.IContext) error { return t.op.before(context, "Locate") }
func (t *Locate) After(context api.IContext) error { return t.op.after() } // frees the executor's ctxs
func (t *Locate) Call(pid int64, it iterator.IReadIterator[string], context api.IContext) ([]string, error) {
	return t.op.call(pid, it, context) // header row only in partition 0 (locate.go:198-204)
}

func NewGrep() any { return &Grep{} }

type Grep struct {
	base.IMapPartitionsWithIndex[string, string]
	op bskOp
}

func (t *Grep) Before(context api.IContext) error { return t.op.before(context, "Grep") }
func (t *Grep) After(context api.IContext) error { return t.op.after() } // frees the executor's ctxs
func (t *Grep) Call(pid int64, it iterator.IReadIterator[string], context api.IContext) ([]string, error) {
	return t.op.call(pid, it, context)
}

// ---- Stats / StatsReduce: bigseqkit-lib/stats.go:16-137.  The map value keeps the reference's encoding
// (length -> count, sentinel keys -1 Q20, -2 Q30, -3 gap sum, -4 alphabet tag) so bigseqkit/stats.go:96-166 is unchanged.
func NewStats() any { return &Stats{} }

type Stats struct {
	base.IMapPartitions[string, map[int64]int64]
	op bskOp
}

func (t *Stats) Before(context api.IContext) error { return t.op.before(context, "Stats") }
func (t *Stats) After(context api.IContext) error { return t.op.after() } // frees the executor's ctxs
func (t *Stats) Call(it iterator.IReadIterator[string], context api.IContext) ([]map[int64]int64, error) {
	ctx := t.op.ctxs[context.ThreadId()]
	C.bsk_reset(ctx)
	if _, err := t.op.call(0, it, context); err != nil {
		return nil, err
	}
	var s C.bsk_stats
	if rc := C.bsk_stats_result(ctx, &s); rc != C.BSK_OK {
		return nil, fmt.Errorf("%s", C.GoString(C.bsk_last_error(ctx)))
	}
	m := make(map[int64]int64, int(s.n_hist)+4)
	lens := unsafe.Slice((*uint64)(unsafe.Pointer(s.hist_len)), int(s.n_hist))
	cnts := unsafe.Slice((*uint64)(unsafe.Pointer(s.hist_cnt)), int(s.n_hist))
	for i := range lens {
		m[int64(lens[i])] = int64(cnts[i])
	}
	m[-1], m[-2], m[-3] = int64(s.q20), int64(s.q30), int64(s.sum_gap)
	switch C.GoString(&s._type[0]) {
	case "DNA":
		m[-4] = 'D'
	case "RNA":
		m[-4] = 'R'
	case "":
		m[-4] = 'U'
	default:
		m[-4] = 'F'
	}
	return []map[int64]int64{m}, nil
}

// ---- StatsReduce: bigseqkit-lib/stats.go:119-137, the function bigseqkit/stats.go:91 hands to Reduce.  Sum
// semantics (SURVEY Q2: the snapshot's `result[k] = v` drops counts as soon as there are two partitions); the
// alphabet tag -4 keeps the first partition's value.
func NewStatsReduce() any { return &StatsReduce{} }

type StatsReduce struct {
	base.IReduce[map[int64]int64]
	base.IOnlyCall
}

func (t *StatsReduce) Call(v1 map[int64]int64, v2 map[int64]int64, context api.IContext) (map[int64]int64, error) {
	for k, v := range v2 {
		if k == -4 {
			if _, ok := v1[k]; !ok {
				v1[k] = v
			}
			continue
		}
		v1[k] += v
	}
	return v1, nil
}

// ---- RmDupPrepare: bigseqkit-lib/rmdup.go:23-90.  Same element type as the reference --
// IPair[int64(xxhash.Sum64(subject)), Record.Format(LineWidth)] -- so bigseqkit/rmdup.go:92-108 (GroupByKey, then the
// reference's own RmDupCheck) runs unchanged.  One bsk_run_buffer call per partition parses, hashes (XXH64 on the
// device, bit-exact with cespare/xxhash) and formats; bsk_rmdup_keys returns the keys in record order.
func NewRmDupPrepare() any { return &RmDupPrepare{} }

type RmDupPrepare struct {
	base.IMapPartitions[string, ipair.IPair[int64, string]]
	op bskOp
}

func (t *RmDupPrepare) Before(context api.IContext) error { return t.op.before(context, "RmDupPrepare") }
func (t *RmDupPrepare) After(context api.IContext) error  { return t.op.after() }
func (t *RmDupPrepare) Call(it iterator.IReadIterator[string], context api.IContext) ([]ipair.IPair[int64, string], error) {
	vals, err := t.op.call(0, it, context)
	if err != nil {
		return nil, err
	}
	ctx := t.op.ctxs[context.ThreadId()]
	var kp *C.int64_t
	var kn C.size_t
	if rc := C.bsk_rmdup_keys(ctx, &kp, &kn); rc != C.BSK_OK {
		return nil, fmt.Errorf("%s", C.GoString(C.bsk_last_error(ctx)))
	}
	if int(kn) != len(vals) {
		return nil, fmt.Errorf("bigseqkit-b200: %d keys for %d records", int(kn), len(vals))
	}
	keys := unsafe.Slice((*int64)(unsafe.Pointer(kp)), int(kn))
	res := make([]ipair.IPair[int64, string], len(vals))
	for i := range vals {
		res[i] = *ipair.New(keys[i], vals[i]+"\n") // Format() ends in '\n' (rmdup.go:86); RmDupCheck re-parses it
	}
	return res, nil
}

// ---- RmDupSharded (optional, needs the driver patch of INTEGRATION.md): the whole rmdup of bigseqkit/rmdup.go:92-108
// as ONE MapPartitions.  Every executor hashes its partition on its GPU, the executors exchange 16-byte fingerprints
// with one NCCL all-gather (bsk_rmdup_sharded) and each keeps the records whose subject was not seen earlier in
// global input order -- instead of moving every record through the GroupByKey shuffle.  One partition per executor
// (the driver repartitions to Executors()); the NCCL unique id travels through the executors' MPI group.
func NewRmDupSharded() any { return &RmDup{} }

type RmDup struct {
	base.IMapPartitions[string, string]
	op bskOp
}

func (t *RmDup) Before(context api.IContext) error {
	if err := t.op.before(context, "RmDup"); err != nil {
		return err
	}
	id := make([]byte, C.BSK_COMM_ID_BYTES)
	if context.ExecutorId() == 0 {
		if rc := C.bsk_comm_unique_id((*C.uint8_t)(unsafe.Pointer(&id[0]))); rc != C.BSK_OK {
			return fmt.Errorf("%s", C.GoString(C.bsk_comm_error()))
		}
	}
	if err := context.MpiGroup().Bcast(id, 0); err != nil { // 128 bytes from executor 0 to all
		return err
	}
	if rc := C.bsk_comm_init(t.op.ctxs[0], (*C.uint8_t)(unsafe.Pointer(&id[0])), C.int(context.Executors()), C.int(context.ExecutorId())); rc != C.BSK_OK {
		return fmt.Errorf("%s", C.GoString(C.bsk_last_error(t.op.ctxs[0])))
	}
	return nil
}
func (t *RmDup) Call(it iterator.IReadIterator[string], context api.IContext) ([]string, error) {
	return t.op.callSharded(it, context)
}

// After writes the -d / -D files per executor like bigseqkit-lib/rmdup.go:245-275 (flag meaning, not the
// directory swap of :246-267): <DupSeqsFile>/<executor id> and <DupNumFile>/<executor id>, the texts of the
// executor's ctxs one after the other.
func (t *RmDup) After(context api.IContext) error {
	opts := bigseqkit.StringToOptions[bigseqkit.RmDupOptions](context.Vars()["opts"].(string))
	write := func(dir string, numbers bool) error {
		if len(dir) == 0 {
			return nil
		}
		var text []byte
		for _, c := range t.op.ctxs {
			var p *C.char
			var n C.size_t
			var rc C.int
			if numbers {
				rc = C.bsk_rmdup_dup_num(c, &p, &n)
			} else {
				rc = C.bsk_rmdup_dup_seqs(c, &p, &n)
			}
			if rc != C.BSK_OK {
				return fmt.Errorf("%s", C.GoString(C.bsk_last_error(c)))
			}
			text = append(text, C.GoBytes(unsafe.Pointer(p), C.int(n))...)
		}
		if len(text) == 0 {
			return nil
		}
		if err := os.MkdirAll(dir, os.ModePerm); err != nil {
			return err
		}
		return os.WriteFile(path.Join(dir, strconv.Itoa(context.ExecutorId())), text, 0o644)
	}
	if err := write(*opts.DupSeqsFile, false); err != nil {
		return err
	}
	if err := write(*opts.DupNumFile, true); err != nil {
		return err
	}
	return t.op.after()
}
