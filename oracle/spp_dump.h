/*
 * spp_dump.h -- tiny binary container used by the oracle tools (TEST INFRASTRUCTURE ONLY).
 *
 * A dump file is a sequence of records:
 *     char     name[32]   (zero padded)
 *     uint64_t dtype      (0 = float64, 1 = uint64)
 *     uint64_t count
 *     payload  count * 8 bytes
 * The Python reader lives in slam_plus_plus_b200/sppio.py.
 *
 * A graph file ("SPPGRAF1") is described in slam_plus_plus_b200/sppio.py as well; the
 * reader below is what the reference driver and the C oracle use.
 */
#ifndef SPP_DUMP_H
#define SPP_DUMP_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

static inline int spp_dump_record(FILE *f, const char *name, uint64_t dtype, uint64_t count, const void *data)
{
	char nm[32];
	memset(nm, 0, sizeof(nm));
	strncpy(nm, name, 31);
	if(fwrite(nm, 1, 32, f) != 32) return -1;
	if(fwrite(&dtype, 8, 1, f) != 1) return -1;
	if(fwrite(&count, 8, 1, f) != 1) return -1;
	if(count && fwrite(data, 8, count, f) != count) return -1;
	return 0;
}

static inline int spp_dump_f64(FILE *f, const char *name, uint64_t count, const double *data)
{
	return spp_dump_record(f, name, 0, count, data);
}

static inline int spp_dump_u64(FILE *f, const char *name, uint64_t count, const uint64_t *data)
{
	return spp_dump_record(f, name, 1, count, data);
}

/* graph kinds */
enum { SPP_GRAPH_BA = 0, SPP_GRAPH_SE2 = 1, SPP_GRAPH_SE3 = 2 };

typedef struct {
	uint64_t kind, n_vertices, n_edges;
	/* BA: vtype[v] 0 = camera (11 doubles: t, axis-angle, fx fy cx cy d), 1 = point (3 doubles) */
	uint64_t *vtype;   /* n_vertices (BA only) */
	uint64_t *voff;    /* n_vertices + 1 offsets into vdata */
	double *vdata;
	uint64_t *e0, *e1; /* BA: e0 = point vertex id, e1 = camera vertex id; pose graphs: from, to */
	double *z;         /* measurement, zdim per edge */
	double *info;      /* information matrix, zdim*zdim per edge (symmetric, full) */
	uint64_t zdim;
} spp_graph_t;

static inline void spp_graph_free(spp_graph_t *g)
{
	free(g->vtype); free(g->voff); free(g->vdata); free(g->e0); free(g->e1); free(g->z); free(g->info);
	memset(g, 0, sizeof(*g));
}

static inline int spp_graph_read(const char *path, spp_graph_t *g)
{
	memset(g, 0, sizeof(*g));
	FILE *f = fopen(path, "rb");
	if(!f) return -1;
	char magic[8];
	uint64_t hdr[3];
	if(fread(magic, 1, 8, f) != 8 || memcmp(magic, "SPPGRAF1", 8) || fread(hdr, 8, 3, f) != 3) {
		fclose(f);
		return -2;
	}
	g->kind = hdr[0]; g->n_vertices = hdr[1]; g->n_edges = hdr[2];
	const uint64_t nv = g->n_vertices, ne = g->n_edges;
	g->voff = (uint64_t*)malloc((nv + 1) * 8);
	g->vtype = (uint64_t*)malloc((nv ? nv : 1) * 8);
	int ok = 1;
	if(g->kind == SPP_GRAPH_BA) {
		ok = ok && fread(g->vtype, 8, nv, f) == nv;
		g->zdim = 2;
	} else {
		for(uint64_t i = 0; i < nv; ++ i) g->vtype[i] = 0;
		g->zdim = (g->kind == SPP_GRAPH_SE2)? 3 : 6;
	}
	g->voff[0] = 0;
	for(uint64_t i = 0; ok && i < nv; ++ i) {
		uint64_t d = (g->kind == SPP_GRAPH_BA)? ((g->vtype[i] == 0)? 11 : 3) : g->zdim;
		g->voff[i + 1] = g->voff[i] + d;
	}
	const uint64_t nvd = ok? g->voff[nv] : 0, zd = g->zdim;
	g->vdata = (double*)malloc((nvd ? nvd : 1) * 8);
	g->e0 = (uint64_t*)malloc((ne ? ne : 1) * 8);
	g->e1 = (uint64_t*)malloc((ne ? ne : 1) * 8);
	g->z = (double*)malloc((ne ? ne : 1) * zd * 8);
	g->info = (double*)malloc((ne ? ne : 1) * zd * zd * 8);
	ok = ok && fread(g->vdata, 8, nvd, f) == nvd;
	ok = ok && fread(g->e0, 8, ne, f) == ne;
	ok = ok && fread(g->e1, 8, ne, f) == ne;
	ok = ok && fread(g->z, 8, ne * zd, f) == ne * zd;
	ok = ok && fread(g->info, 8, ne * zd * zd, f) == ne * zd * zd;
	fclose(f);
	if(!ok) {
		spp_graph_free(g);
		return -3;
	}
	return 0;
}

#ifdef __cplusplus
}
#endif

#endif /* SPP_DUMP_H */
