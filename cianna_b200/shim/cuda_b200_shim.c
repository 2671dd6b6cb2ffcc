/*
 * cuda_b200_shim.c - the upstream-side binding of the B200 core: every `cuda_*` symbol upstream CIANNA's host code binds
 * (src/prototypes.h:217-295 upstream) implemented over libcianna_host.so / libcianna_b200.so.
 *
 * HOW A MAINTAINER USES IT: build upstream's host sources with -D CUDA exactly as its own compile script does, but link
 * this one C file (gcc, no nvcc) plus -lcianna_host -lcianna_b200 in place of the seven src/cuda/ *.cu objects, cuBLAS
 * and cuRAND.  Nothing in upstream's C or Python sources changes: `cnn.init(..., comp_meth="C_CUDA")` then runs the
 * sm_100a kernels.  oracle/build_ref.sh does exactly that (variant "dropin", tests/test_gpu_backends.py drives it).
 *
 * This file includes upstream's OWN headers (prototypes.h -> structs.h): it sees upstream's `network`, `layer`,
 * `conv_param`... and therefore cannot include cianna_b200/host/cianna.h (same names); everything it needs from the
 * host library goes through the plain-argument cbb_* calls of cianna_bridge.h.
 *
 * The principle: for every upstream layer handed to cuda_convert_X_layer, the same layer is created in a MIRROR
 * network inside libcianna_host.so (upstream's geometry, activation string and initial weights); upstream's big host
 * staging arrays (im2col buffers, rotated filters...) are released, and the pointers upstream's host code keeps using
 * (layer->output, FP32_filters, update, norm tables, IoU_monitor) are pointed at the mirror's device memory.
 * upstream's layer->forward / layer->backprop then run the mirrored layer's kernels.
 *
 * Differences a caller can observe, all by design:
 *   - activations live in the core's channels-last layout; cuda_get_table_to_FP32 recognises a layer's output /
 *     delta_o pointer and hands back upstream's [C][B][H*W] (dense: [B][n+1]) layout, so auxil.c reads what it expects;
 *   - conv_param->TC_padding is reset to 0 (the core pads its own operand copies; FP32_filters rows are flat_f_size
 *     long, which is also the save-file layout);
 *   - `update` (momentum) tables are FP32 in every mode;
 *   - TF32C_FP32A and FP16C_FP16A are refused with an error (the core provides FP32, FP16C_FP32A, BF16C_FP32A).
 */
#include "prototypes.h"
#include "cianna_b200.h"
#include "cianna_bridge.h"

#define SHIM_CHECK(call) do { int rc_ = (call); if (rc_ != CB200_OK) { \
	printf("\nERROR: cianna_b200 shim: %s failed (%d): %s\n", #call, rc_, cb200_last_error()); exit(EXIT_FAILURE); } } while (0)

/* ------------------------------------------------------------------ mirror bookkeeping */
enum { R_ACT = 1, R_F32 = 2, R_MASTER = 3 };
typedef struct { const char *ptr; size_t bytes; int kind, net, layer, what; } region;
static region *regions = NULL;
static int nb_regions = 0, cap_regions = 0;
static int mirror_ready[MAX_NETWORKS_NB];
static float *norm_shadow[MAX_NETWORKS_NB][MAX_LAYERS_NB];   /* gamma | beta as last seen on both sides */

static void region_add(const void *ptr, size_t bytes, int kind, int net, int l, int what)
{
	if (ptr == NULL) return;
	if (nb_regions == cap_regions) {
		cap_regions = cap_regions ? 2 * cap_regions : 256;
		regions = (region *)realloc(regions, cap_regions * sizeof(region));
	}
	regions[nb_regions].ptr = (const char *)ptr; regions[nb_regions].bytes = bytes; regions[nb_regions].kind = kind;
	regions[nb_regions].net = net; regions[nb_regions].layer = l; regions[nb_regions].what = what;
	nb_regions++;
}

static region *region_find(const void *ptr)
{
	int i;
	for (i = nb_regions - 1; i >= 0; i--)
		if ((const char *)ptr >= regions[i].ptr && (const char *)ptr < regions[i].ptr + regions[i].bytes) return &regions[i];
	return NULL;
}

static void regions_drop_net(int net)
{
	int i, n = 0;
	for (i = 0; i < nb_regions; i++) if (regions[i].net != net) regions[n++] = regions[i];
	nb_regions = n;
}

static int dtype_of(network *net)
{
	switch (net->cu_inst.use_cuda_TC) {
	case FP16C_FP32A: case FP16C_FP16A: return CB200_FP16;
	case BF16C_FP32A: return CB200_BF16;
	default: return CB200_FP32;
	}
}
static size_t esize(network *net) { return cb200_dtype_size(dtype_of(net)); }

static int index_of(layer *cur)
{
	network *net = cur->c_network;
	int i;
	for (i = net->nb_layers - 1; i >= 0; i--) if (net->net_layers[i] == cur) return i;
	printf("\nERROR: cianna_b200 shim: layer not found in its network\n");
	exit(EXIT_FAILURE);
}

/* init_cuda runs before init_network has stored the dimensions (src/auxil.c:164-200 upstream): the mirror network is
 * created at the first call that needs it */
static int ensure_mirror(network *net)
{
	if (!mirror_ready[net->id]) {
		int dims[4] = { net->in_dims[0], net->in_dims[1], net->in_dims[2], net->in_dims[3] };
		int l;
		regions_drop_net(net->id);
		for (l = 0; l < MAX_LAYERS_NB; l++) { free(norm_shadow[net->id][l]); norm_shadow[net->id][l] = NULL; }
		cbb_init(net->id, dims, net->output_dim, net->input_bias, net->batch_param == SGD ? 1 : net->batch_size,
			net->cu_inst.dynamic_load, net->cu_inst.use_cuda_TC, net->inference_only, net->adv_size);
		mirror_ready[net->id] = 1;
	}
	return net->id;
}

/* upstream's pointer for a mirrored activation tensor: the mirror's own buffer, or (a group-norm layer, which may be
 * evaluated inside the following pooling kernel and then has none) a 256-byte token allocation that only serves as a key
 * of the region table */
static void *act_handle(network *net, int l, int want_delta)
{
	void *p = net->net_layers[l]->type == NORM ? NULL : cbb_act_ptr(net->id, l, want_delta);
	int chw[3];
	size_t bytes;
	if (want_delta && net->inference_only) return NULL;
	cbb_layer_shape(net->id, l, chw);
	bytes = (size_t)net->batch_size * chw[1] * chw[2] * cb200_round_channels(chw[0]) * esize(net);
	/* (a norm layer's own buffers are released when the NEXT layer turns out to be a pooling layer that absorbs it:
	 * never hand them out) */
	if (p == NULL) { SHIM_CHECK(cb200_malloc(&p, 256)); bytes = 256; }
	region_add(p, bytes, R_ACT, net->id, l, want_delta);
	return p;
}

static void free_host(void **p) { if (*p != NULL) { free(*p); *p = NULL; } }

/* ------------------------------------------------------------------ set-up */
void init_cuda(network *net)
{
	if (net->cu_inst.use_cuda_TC == TF32C_FP32A || net->cu_inst.use_cuda_TC == FP16C_FP16A) {
		printf("\nERROR: the B200 core provides FP32C_FP32A, FP16C_FP32A and BF16C_FP32A; TF32C_FP32A / FP16C_FP16A are not available.\n");
		exit(EXIT_FAILURE);
	}
	SHIM_CHECK(cb200_init(-1));
	mirror_ready[net->id] = 0;      /* (re-)initialised network: a fresh mirror at the first layer / dataset call */
	is_cuda_init = 1;
}

void cuda_set_TC_scale_factor(network *net, float val)
{
	static int warned = 0;
	if (net->cu_inst.use_cuda_TC == FP16C_FP32A || net->cu_inst.use_cuda_TC == FP16C_FP16A)
		net->TC_scale_factor = val;
	else {
		if (!warned && val != 1.0f)
			printf("\nWARNING: Tried to set TC_scale_factor but the mixed precision mode is incompatible.\nScale kept to 1.\n");
		warned = 1;
		net->TC_scale_factor = 1.0f;
	}
}

void cuda_sync(void) { SHIM_CHECK(cb200_device_sync()); }

/* ------------------------------------------------------------------ tables */
void cuda_free_table(void *tab)
{
	region *r = region_find(tab);
	if (tab == NULL || r != NULL) return;      /* mirror-owned memory goes with the mirror */
	SHIM_CHECK(cb200_free(tab));
}

void cuda_create_host_table(network *net, void **tab, size_t size) { *tab = malloc(size * esize(net)); }

static void upload(void *dev, const void *host, size_t bytes)
{
	SHIM_CHECK(cb200_h2d(dev, host, bytes, NULL));
	SHIM_CHECK(cb200_stream_sync(NULL));
}

static void download(void *host, const void *dev, size_t bytes)
{
	SHIM_CHECK(cb200_d2h(host, dev, bytes, NULL));
	SHIM_CHECK(cb200_stream_sync(NULL));
}

size_t cuda_convert_table(network *net, void **tab, size_t size, int keep_host)
{
	size_t es = esize(net);
	void *host = *tab, *typed = host, *dev = NULL;
	if (es != 4) {
		typed = malloc(size * es);
		SHIM_CHECK(cb200_host_cast_from_f32(typed, dtype_of(net), (const float *)host, size));
		free(host);
	}
	SHIM_CHECK(cb200_malloc(&dev, size * es));
	upload(dev, typed, size * es);
	if (keep_host == 0) free(typed);
	*tab = dev;
	return size * es;
}

size_t cuda_convert_table_FP32(void **tab, size_t size, int keep_host)
{
	void *host = *tab, *dev = NULL;
	SHIM_CHECK(cb200_malloc(&dev, size * sizeof(float)));
	upload(dev, host, size * sizeof(float));
	if (keep_host == 0) free(host);
	*tab = dev;
	return size * sizeof(float);
}

size_t cuda_convert_table_int(int **tab, size_t size, int keep_host)
{
	int *host = *tab;
	void *dev = NULL;
	SHIM_CHECK(cb200_malloc(&dev, size * sizeof(int)));
	upload(dev, host, size * sizeof(int));
	if (keep_host == 0) free(host);
	*tab = (int *)dev;
	return size * sizeof(int);
}

void cuda_create_table_FP32(void **tab, size_t size) { SHIM_CHECK(cb200_malloc(tab, size * sizeof(float))); }
void cuda_create_table(network *net, void **tab, size_t size) { SHIM_CHECK(cb200_malloc(tab, size * esize(net))); }

void cuda_set_mem_value(void *device_mem_loc, float value, size_t size)
{
	/* upstream's only use: the pivot weight of the dense layer below a new dense layer (src/dense_layer.c:209-221);
	 * the mirror's dense_create has set it already, in the right table for every precision mode */
	if (region_find(device_mem_loc) != NULL) return;
	upload(device_mem_loc, &value, size);
}

void cuda_master_weight_copy(network *net, float *master, void *copy, size_t size)
{
	if ((void *)master == copy) return;
	SHIM_CHECK(cb200_cast_from_f32(copy, dtype_of(net), master, size, NULL));
}

void cuda_get_table_FP32(void *cuda_table, void *table, size_t size) { download(table, cuda_table, size * sizeof(float)); }

void cuda_get_table_to_FP32(network *net, void *cuda_table, float *table, size_t size, void *buffer)
{
	region *r = region_find(cuda_table);
	size_t es = esize(net);
	if (r != NULL && r->kind == R_ACT) {
		cbb_export_act(r->net, r->layer, r->what, table);      /* upstream's layout, whole tensor */
		return;
	}
	if ((r != NULL && r->kind != R_ACT) || es == 4) { download(table, cuda_table, size * sizeof(float)); return; }
	{
		void *tmp = buffer != NULL ? buffer : malloc(size * es);
		download(tmp, cuda_table, size * es);
		SHIM_CHECK(cb200_host_cast_to_f32(table, tmp, dtype_of(net), size));
		if (buffer == NULL) free(tmp);
	}
}

void cuda_get_table(network *net, void *cuda_table, void *table, size_t size) { download(table, cuda_table, size * esize(net)); }

void cuda_put_table_FP32(void *cuda_table, void *table, size_t size)
{
	region *r = region_find(cuda_table);
	upload(cuda_table, table, size * sizeof(float));
	if (r != NULL && r->kind == R_MASTER) cbb_weights_changed(r->net, r->layer);
}

void cuda_get_typed_host_table(network *net, void *typed_table, float *out_table, size_t size)
{
	SHIM_CHECK(cb200_host_cast_to_f32(out_table, typed_table, dtype_of(net), size));
}

void cuda_put_table(network *net, void *cuda_table, void *table, size_t size) { upload(cuda_table, table, size * esize(net)); }

static void print_floats(const float *v, size_t size, int return_every)
{
	size_t i;
	for (i = 0; i < size; i++) {
		if (i % return_every == 0) printf("\n");
		printf("%g \t ", v[i]);
	}
}

void cuda_print_table_FP32(void *tab, size_t size, int return_every)
{
	float *tmp = (float *)malloc(size * sizeof(float));
	download(tmp, tab, size * sizeof(float));
	print_floats(tmp, size, return_every);
	free(tmp);
}

void cuda_print_table(network *net, void *tab, size_t size, int return_every)
{
	float *tmp = (float *)malloc(size * sizeof(float));
	cuda_get_table_to_FP32(net, tab, tmp, size, NULL);
	print_floats(tmp, size, return_every);
	printf("\n");
	free(tmp);
}

void cuda_print_table_int(network *net, int *tab, size_t size, int return_every)
{
	int *tmp = (int *)malloc(size * sizeof(int));
	size_t i;
	(void)net;
	download(tmp, tab, size * sizeof(int));
	printf("\n");
	for (i = 0; i < size; i++) {
		if (i % return_every == 0) printf("\n");
		printf("%d ", tmp[i]);
	}
	printf("\n");
	free(tmp);
}

void cuda_random_vector(float *tab, size_t size)
{
	/* only upstream's own CUDA layers call this (dropout masks); the core draws its masks inside its kernels.  Kept for
	 * completeness: uniform [0,1) values drawn on the host */
	float *tmp = (float *)malloc(size * sizeof(float));
	size_t i;
	for (i = 0; i < size; i++) tmp[i] = (float)(rand() / ((double)RAND_MAX + 1.0));
	upload(tab, tmp, size * sizeof(float));
	free(tmp);
}

/* ------------------------------------------------------------------ datasets */
static void copy_to_typed_FP32(float *in, void *out, int out_offset, size_t n) { memcpy((float *)out + out_offset, in, n * sizeof(float)); }
static void copy_to_typed_FP16(float *in, void *out, int out_offset, size_t n) { cb200_host_cast_from_f32((uint16_t *)out + out_offset, CB200_FP16, in, n); }
static void copy_to_typed_BF16(float *in, void *out, int out_offset, size_t n) { cb200_host_cast_from_f32((uint16_t *)out + out_offset, CB200_BF16, in, n); }

Dataset cuda_create_dataset(network *net, int nb_elem)
{
	Dataset data;
	size_t es = esize(net), in_row = net->input_dim + 1;
	int i, j, dt = dtype_of(net);
	float bias = net->input_bias;

	memset(&data, 0, sizeof(data));
	data.size = nb_elem;
	data.nb_batch = (data.size - 1) / net->batch_size + 1;
	data.input = (void **)malloc(data.nb_batch * sizeof(void *));
	data.target = (void **)malloc(data.nb_batch * sizeof(void *));
	data.localization = HOST;
	data.cont_copy = dt == CB200_FP32 ? copy_to_typed_FP32 : (dt == CB200_FP16 ? copy_to_typed_FP16 : copy_to_typed_BF16);
	for (i = 0; i < data.nb_batch; i++) {
		data.input[i] = calloc((size_t)net->batch_size * in_row, es);
		data.target[i] = calloc((size_t)net->batch_size * net->output_dim, es);
		for (j = 0; j < net->batch_size; j++)
			data.cont_copy(&bias, data.input[i], (int)(j * in_row + net->input_dim), 1);
	}
	return data;
}

static void publish_batch_pointers(Dataset *data)
{
	SHIM_CHECK(cb200_malloc((void **)&data->input_device, data->nb_batch * sizeof(void *)));
	SHIM_CHECK(cb200_malloc((void **)&data->target_device, data->nb_batch * sizeof(void *)));
	upload(data->input_device, data->input, data->nb_batch * sizeof(void *));
	upload(data->target_device, data->target, data->nb_batch * sizeof(void *));
	data->localization = DEVICE;
}

/* typed host batches (cuda_create_dataset + cont_copy) -> device (src/cuda/cuda_main.cu:397-406 upstream) */
void cuda_get_batched_dataset(network *net, Dataset *data)
{
	size_t es = esize(net), n_in = (size_t)net->batch_size * (net->input_dim + 1), n_out = (size_t)net->batch_size * net->output_dim;
	int i;
	for (i = 0; i < data->nb_batch; i++) {
		void *host_in = data->input[i], *host_tg = data->target[i];
		SHIM_CHECK(cb200_malloc(&data->input[i], n_in * es));
		SHIM_CHECK(cb200_malloc(&data->target[i], (n_out ? n_out : 1) * es));
		SHIM_CHECK(cb200_h2d(data->input[i], host_in, n_in * es, NULL));
		SHIM_CHECK(cb200_h2d(data->target[i], host_tg, n_out * es, NULL));
		SHIM_CHECK(cb200_stream_sync(NULL));
		free(host_in); free(host_tg);
	}
	publish_batch_pointers(data);
}

/* FP32 host batches -> typed device batches.  upstream converts on the host, one element at a time
 * (src/cuda/cuda_main.cu:355-371); here the FP32 batch is staged on the device and converted there with the same
 * round-toward-zero rule */
void cuda_convert_dataset(network *net, Dataset *data)
{
	size_t es = esize(net), n_in = (size_t)net->batch_size * (net->input_dim + 1), n_out = (size_t)net->batch_size * net->output_dim;
	size_t n_max = n_in > n_out ? n_in : n_out;
	float *stage = NULL;
	int i, dt = dtype_of(net);
	SHIM_CHECK(cb200_malloc((void **)&stage, n_max * sizeof(float)));
	for (i = 0; i < data->nb_batch; i++) {
		void *host_in = data->input[i], *host_tg = data->target[i];
		SHIM_CHECK(cb200_malloc(&data->input[i], n_in * es));
		SHIM_CHECK(cb200_malloc(&data->target[i], (n_out ? n_out : 1) * es));
		SHIM_CHECK(cb200_h2d(stage, host_in, n_in * sizeof(float), NULL));
		SHIM_CHECK(cb200_cast_from_f32_rz(data->input[i], dt, stage, n_in, NULL));
		SHIM_CHECK(cb200_h2d(stage, host_tg, n_out * sizeof(float), NULL));
		SHIM_CHECK(cb200_cast_from_f32_rz(data->target[i], dt, stage, n_out, NULL));
		SHIM_CHECK(cb200_stream_sync(NULL));
		free(host_in); free(host_tg);
	}
	SHIM_CHECK(cb200_free(stage));
	publish_batch_pointers(data);
}

/* FP32 host batches -> typed host batches; like upstream only the INPUT table is converted (src/cuda/cuda_main.cu:437-442) */
void cuda_convert_host_dataset(network *net, Dataset *data)
{
	size_t es = esize(net), n_in = (size_t)net->batch_size * (net->input_dim + 1);
	int i;
	if (es != 4)
		for (i = 0; i < data->nb_batch; i++) {
			void *typed = malloc(n_in * es);
			SHIM_CHECK(cb200_host_cast_from_f32(typed, dtype_of(net), (const float *)data->input[i], n_in));
			free(data->input[i]);
			data->input[i] = typed;
		}
	data->localization = HOST;
}

void cuda_free_dataset(Dataset *data)
{
	int i;
	if (data->input != NULL)
		for (i = 0; i < data->nb_batch; i++) { cb200_free(data->input[i]); cb200_free(data->target[i]); }
	if (data->input_device != NULL) { cb200_free(data->input_device); cb200_free(data->target_device); }
}

/* ------------------------------------------------------------------ shuffles (src/cuda/cuda_main.cu:590-760 upstream) */
void cuda_shuffle(network *net, Dataset data, Dataset duplicate, int *index_shuffle, int *index_shuffle_device)
{
	size_t es = esize(net);
	int i, j, temp;
	for (i = 0; i < data.size - 1; i++) {      /* the same draw sequence as upstream */
		j = i + (int)((rand() / ((double)RAND_MAX)) * (double)(data.size - i));
		temp = index_shuffle[i]; index_shuffle[i] = index_shuffle[j]; index_shuffle[j] = temp;
	}
	SHIM_CHECK(cb200_h2d(index_shuffle_device, index_shuffle, data.size * sizeof(int), NULL));
	SHIM_CHECK(cb200_rows_permute(duplicate.input_device, data.input_device, index_shuffle_device, data.size, net->batch_size, (net->input_dim + 1) * es, NULL));
	SHIM_CHECK(cb200_rows_permute(duplicate.target_device, data.target_device, index_shuffle_device, data.size, net->batch_size, net->output_dim * es, NULL));
	SHIM_CHECK(cb200_rows_permute(data.input_device, duplicate.input_device, NULL, data.size, net->batch_size, (net->input_dim + 1) * es, NULL));
	SHIM_CHECK(cb200_rows_permute(data.target_device, duplicate.target_device, NULL, data.size, net->batch_size, net->output_dim * es, NULL));
	SHIM_CHECK(cb200_stream_sync(NULL));       /* index_shuffle is pageable host memory */
}

static void swap_rows(void *a, void *b, size_t bytes, void *tmp)
{
	memcpy(tmp, a, bytes); memcpy(a, b, bytes); memcpy(b, tmp, bytes);
}

static void host_fisher_yates(network *net, void **input, void **target, int size)
{
	size_t es = esize(net), in_bytes = (net->input_dim + 1) * es, out_bytes = net->output_dim * es;
	void *tmp = malloc(in_bytes > out_bytes ? in_bytes : out_bytes);
	int i, j;
	for (i = 0; i < size - 1; i++) {
		j = i + (int)((rand() / ((double)RAND_MAX)) * (double)(size - i));
		swap_rows((char *)input[i / net->batch_size] + (size_t)(i % net->batch_size) * in_bytes,
			(char *)input[j / net->batch_size] + (size_t)(j % net->batch_size) * in_bytes, in_bytes, tmp);
		swap_rows((char *)target[i / net->batch_size] + (size_t)(i % net->batch_size) * out_bytes,
			(char *)target[j / net->batch_size] + (size_t)(j % net->batch_size) * out_bytes, out_bytes, tmp);
	}
	free(tmp);
}

void cuda_host_shuffle(network *net, Dataset data, Dataset duplicate)
{
	size_t es = esize(net), n_in = (size_t)net->batch_size * (net->input_dim + 1) * es, n_out = (size_t)net->batch_size * net->output_dim * es;
	int i;
	for (i = 0; i < data.nb_batch; i++) {
		SHIM_CHECK(cb200_d2h(duplicate.input[i], data.input[i], n_in, NULL));
		SHIM_CHECK(cb200_d2h(duplicate.target[i], data.target[i], n_out, NULL));
	}
	SHIM_CHECK(cb200_stream_sync(NULL));
	host_fisher_yates(net, duplicate.input, duplicate.target, data.size);
	for (i = 0; i < data.nb_batch; i++) {
		SHIM_CHECK(cb200_h2d(data.input[i], duplicate.input[i], n_in, NULL));
		SHIM_CHECK(cb200_h2d(data.target[i], duplicate.target[i], n_out, NULL));
	}
	SHIM_CHECK(cb200_stream_sync(NULL));
}

void cuda_host_only_shuffle(network *net, Dataset data) { host_fisher_yates(net, data.input, data.target, data.size); }

/* ------------------------------------------------------------------ timers (src/cuda/cuda_main.cu:533-588 upstream) */
static void *ev[3][2];
static void timer_init(int k) { if (ev[k][0] == NULL) { SHIM_CHECK(cb200_event_create(&ev[k][0])); SHIM_CHECK(cb200_event_create(&ev[k][1])); } }
static void timer_in(int k) { timer_init(k); SHIM_CHECK(cb200_event_record(ev[k][0], NULL)); }
static float timer_out(int k)
{
	float ms = 0.0f;
	timer_init(k);
	SHIM_CHECK(cb200_event_record(ev[k][1], NULL));
	SHIM_CHECK(cb200_event_elapsed_ms(ev[k][0], ev[k][1], &ms));
	return ms * 1000.0f;      /* microseconds */
}
void cuda_perf_eval_init(void) { timer_init(0); }
void cuda_batch_eval_init(void) { timer_init(1); }
void cuda_epoch_eval_init(void) { timer_init(2); }
void cuda_perf_eval_in(void) { timer_in(0); }
void cuda_batch_eval_in(void) { timer_in(1); }
void cuda_epoch_eval_in(void) { timer_in(2); }
float cuda_perf_eval_out(void) { return timer_out(0); }
float cuda_batch_eval_out(void) { return timer_out(1); }
float cuda_epoch_eval_out(void) { return timer_out(2); }

/* ------------------------------------------------------------------ layers */
static void norm_push_if_changed(network *net, int l, norm_param *p)
{
	/* upstream keeps gamma / beta in host arrays and uploads them at every pass (src/cuda/cuda_norm_layer.cu:366-367);
	 * the mirror holds them on the device: upload only when the host side was written since the last exchange */
	float *sh = norm_shadow[net->id][l];
	size_t n = p->nb_group;
	if (sh != NULL && memcmp(sh, p->gamma, n * sizeof(float)) == 0 && memcmp(sh + n, p->beta, n * sizeof(float)) == 0) return;
	if (sh == NULL) sh = norm_shadow[net->id][l] = (float *)malloc(2 * n * sizeof(float));
	memcpy(sh, p->gamma, n * sizeof(float)); memcpy(sh + n, p->beta, n * sizeof(float));
	cbb_norm_set(net->id, l, p->gamma, p->beta);
}

static void b200_forward(layer *cur)
{
	network *net = cur->c_network;
	int l = index_of(cur);
	if (net->length == 0) return;
	if (cur->type == NORM) norm_push_if_changed(net, l, (norm_param *)cur->param);
	cbb_forward_layer(net->id, l, net->input, net->length, net->is_inference, net->inference_drop_mode == MC_MODEL);
}

static void b200_backprop(layer *cur)
{
	network *net = cur->c_network;
	int l = index_of(cur), k;
	cbb_backprop_layer(net->id, l, net->learning_rate, net->momentum, net->weight_decay, cur->frozen);
	if (l != 0) return;
	/* the optimizer sweep has run: bring the (tiny) group-norm parameters back to upstream's host arrays, which
	 * norm_save and the next pass read */
	for (k = 0; k < net->nb_layers; k++)
		if (net->net_layers[k]->type == NORM) {
			norm_param *p = (norm_param *)net->net_layers[k]->param;
			cbb_norm_get_async(net->id, k, p->gamma, p->beta);
		}
	cbb_stream_sync();
	for (k = 0; k < net->nb_layers; k++)
		if (net->net_layers[k]->type == NORM && norm_shadow[net->id][k] != NULL) {
			norm_param *p = (norm_param *)net->net_layers[k]->param;
			memcpy(norm_shadow[net->id][k], p->gamma, p->nb_group * sizeof(float));
			memcpy(norm_shadow[net->id][k] + p->nb_group, p->beta, p->nb_group * sizeof(float));
		}
}

static void define(layer *cur) { cur->forward = b200_forward; cur->backprop = b200_backprop; }
void cuda_conv_define(layer *current) { define(current); }
void cuda_dense_define(layer *current) { define(current); }
void cuda_pool_define(layer *current) { define(current); }
void cuda_norm_define(layer *current) { define(current); }
void cuda_lrn_define(layer *current) { define(current); }
void cuda_conv_init(network *net) { (void)net; }
void cuda_dense_init(network *net) { (void)net; }
void cuda_pool_init(network *net) { (void)net; }
void cuda_norm_init(network *net) { (void)net; }
void cuda_lrn_init(network *net) { (void)net; }
void init_typed_cuda_activ(network *net) { (void)net; }

static int prev_index(layer *cur) { return cur->previous == NULL ? -1 : index_of(cur->previous); }

static void expect_index(int got, int want)
{
	if (got != want) { printf("\nERROR: cianna_b200 shim: mirror layer %d created for upstream layer %d\n", got, want); exit(EXIT_FAILURE); }
}

static void push_yolo(network *net)
{
	yolo_param *y = net->y_param;
	float slopes[18];
	int i, j;
	for (i = 0; i < 6; i++) for (j = 0; j < 3; j++) slopes[i * 3 + j] = y->slopes_and_maxes_tab[i][j];
	cbb_set_yolo(net->id, y->nb_box, y->nb_class, y->nb_param, y->max_nb_obj_per_image, y->IoU_type, y->prior_dist_type,
		y->prior_size, y->noobj_prob_prior, y->fit_dim, y->strict_box_size_association, y->rand_startup,
		y->rand_prob_best_box_assoc, y->rand_prob, y->min_prior_forced_scaling, y->scale_tab, slopes, y->param_ind_scale,
		y->IoU_limits, y->fit_parts, y->class_softmax, y->diff_flag, y->error_type, y->no_override, y->raw_output);
}

size_t cuda_convert_conv_layer(layer *current)
{
	conv_param *p = (conv_param *)current->param;
	network *net = current->c_network;
	int id = ensure_mirror(net), l = index_of(current);
	size_t nw = (size_t)p->nb_filters * p->flat_f_size;
	char activ[64];

	if (current->activation_type == YOLO) push_yolo(net);
	print_string_activ_param(current, activ);
	expect_index(cbb_conv(id, prev_index(current), p->f_size, p->nb_filters, p->stride, p->padding, p->int_padding, activ,
		current->bias_value, current->dropout_rate, (const float *)p->filters, p->flat_f_size + p->TC_padding), l);

	/* upstream's host staging: not needed any more */
	free_host(&p->filters); free_host(&current->output); free_host(&p->im2col_input);
	if (current->dropout_rate > 0.01f) free_host((void **)&p->dropout_mask);
	if (!net->inference_only) {
		free_host(&p->update); free_host(&p->rotated_filters); free_host(&current->delta_o); free_host(&p->im2col_delta_o);
		if (current->previous != NULL && current->previous->type == DENSE) free_host(&p->temp_delta_o);
	}
	p->TC_padding = 0;
	p->filters = p->FP32_filters = cbb_master(id, l);
	region_add(p->FP32_filters, nw * sizeof(float), R_MASTER, id, l, 0);
	if (!net->inference_only) { p->update = cbb_moment(id, l); region_add(p->update, nw * sizeof(float), R_F32, id, l, 0); }
	current->output = act_handle(net, l, 0);
	current->delta_o = act_handle(net, l, 1);
	return nw * (sizeof(float) + 2 * esize(net));
}

size_t cuda_convert_dense_layer(layer *current)
{
	dense_param *p = (dense_param *)current->param;
	network *net = current->c_network;
	int id = ensure_mirror(net), l = index_of(current);
	size_t nw = (size_t)p->in_size * (p->nb_neurons + 1);
	char activ[64];

	print_string_activ_param(current, activ);
	expect_index(cbb_dense(id, prev_index(current), p->nb_neurons, activ, current->bias_value, current->dropout_rate,
		(const float *)p->weights), l);
	free_host(&p->weights); free_host(&current->output);
	if (current->previous != NULL && current->previous->type != DENSE) {
		free_host(&p->flat_input);
		if (!net->inference_only) free_host(&p->flat_delta_o);
	}
	if (current->dropout_rate > 0.01f) free_host((void **)&p->dropout_mask);
	if (!net->inference_only) { free_host(&p->update); free_host(&current->delta_o); }
	p->weights = p->FP32_weights = cbb_master(id, l);
	region_add(p->FP32_weights, nw * sizeof(float), R_MASTER, id, l, 0);
	if (!net->inference_only) { p->update = cbb_moment(id, l); region_add(p->update, nw * sizeof(float), R_F32, id, l, 0); }
	current->output = act_handle(net, l, 0);
	current->delta_o = act_handle(net, l, 1);
	return nw * (sizeof(float) + 2 * esize(net));
}

size_t cuda_convert_pool_layer(layer *current)
{
	pool_param *p = (pool_param *)current->param;
	network *net = current->c_network;
	int id = ensure_mirror(net), l = index_of(current);
	char activ[64];

	print_string_activ_param(current, activ);
	expect_index(cbb_pool(id, prev_index(current), p->p_size, p->stride, p->padding, p->pool_type == AVG_pool, activ,
		p->global, current->dropout_rate), l);
	free_host(&current->output);
	if (current->dropout_rate > 0.01f) free_host((void **)&p->dropout_mask);
	if (!net->inference_only) { free_host((void **)&p->pool_map); free_host(&current->delta_o); }
	current->output = act_handle(net, l, 0);
	current->delta_o = act_handle(net, l, 1);
	return 0;
}

size_t cuda_convert_norm_layer(layer *current)
{
	norm_param *p = (norm_param *)current->param;
	network *net = current->c_network;
	int id = ensure_mirror(net), l = index_of(current);
	size_t tab = (size_t)p->nb_group * net->batch_size * sizeof(float);
	char activ[64];

	if (current->previous == NULL || current->previous->type == DENSE) {
		printf("\nERROR: group normalization needs a convolutional or pooling layer below it (as upstream: src/norm_layer.c).\n");
		exit(EXIT_FAILURE);
	}
	print_string_activ_param(current, activ);
	expect_index(cbb_norm(id, prev_index(current), activ, p->group_size, p->set_off, p->gamma, p->beta), l);
	norm_shadow[id][l] = (float *)malloc(2 * (size_t)p->nb_group * sizeof(float));
	memcpy(norm_shadow[id][l], p->gamma, p->nb_group * sizeof(float));
	memcpy(norm_shadow[id][l] + p->nb_group, p->beta, p->nb_group * sizeof(float));
	free_host(&current->output); free_host((void **)&p->gamma_gpu); free_host((void **)&p->beta_gpu);
	free_host((void **)&p->mean); free_host((void **)&p->var);
	if (!net->inference_only) { free_host(&current->delta_o); free_host((void **)&p->d_gamma_gpu); free_host((void **)&p->d_beta_gpu); }
	p->gamma_gpu = cbb_norm_table(id, l, 4); p->beta_gpu = cbb_norm_table(id, l, 5);
	p->mean = cbb_norm_table(id, l, 0); p->var = cbb_norm_table(id, l, 1);
	region_add(p->mean, tab, R_F32, id, l, 0); region_add(p->var, tab, R_F32, id, l, 1);
	region_add(p->gamma_gpu, p->nb_group * sizeof(float), R_F32, id, l, 4);
	region_add(p->beta_gpu, p->nb_group * sizeof(float), R_F32, id, l, 5);
	if (!net->inference_only) {
		p->d_gamma_gpu = cbb_norm_table(id, l, 2); p->d_beta_gpu = cbb_norm_table(id, l, 3);
		region_add(p->d_gamma_gpu, tab, R_F32, id, l, 2); region_add(p->d_beta_gpu, tab, R_F32, id, l, 3);
	}
	current->output = act_handle(net, l, 0);
	current->delta_o = act_handle(net, l, 1);
	return 0;
}

size_t cuda_convert_lrn_layer(layer *current)
{
	lrn_param *p = (lrn_param *)current->param;
	network *net = current->c_network;
	int id = ensure_mirror(net), l = index_of(current);
	char activ[64];

	print_string_activ_param(current, activ);
	expect_index(cbb_lrn(id, prev_index(current), activ, p->range, p->k, p->alpha, p->beta), l);
	free_host(&current->output);
	if (!net->inference_only) { free_host(&current->delta_o); free_host((void **)&p->local_scale); }
	current->output = act_handle(net, l, 0);
	current->delta_o = act_handle(net, l, 1);
	return 0;
}

/* ------------------------------------------------------------------ activations and the output error */
static void no_op(layer *current) { (void)current; }      /* activations run inside the mirrored layers' kernels */

void cuda_define_activation(layer *current)
{
	current->activation = no_op;
	current->deriv_activation = no_op;
	if (current->activation_type == YOLO) {
		yolo_param *a = (yolo_param *)current->activ_param;
		conv_param *p = (conv_param *)current->param;
		network *net = current->c_network;
		/* what auxil.c reads after a pass (src/auxil.c:1461 upstream): (objectness, IoU) pairs, -1 where no box was matched */
		a->IoU_monitor = cbb_yolo_monitor(net->id);
		region_add(a->IoU_monitor, 2 * (size_t)a->nb_box * p->nb_area[0] * p->nb_area[1] * p->nb_area[2] * net->batch_size * sizeof(float),
			R_F32, net->id, index_of(current), 9);
	}
}

void cuda_deriv_output_error(layer *current)
{
	network *net = current->c_network;
	cbb_deriv_output_error(net->id, net->target, net->TC_scale_factor, net->iter, net->train.size);
}

void cuda_output_error_fct(layer *current)
{
	network *net = current->c_network;
	cbb_output_error(net->id, net->target, (float *)net->output_error, (size_t)net->batch_size * net->out_size);
}
