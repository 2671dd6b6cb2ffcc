/*
 * network.c - network object, datasets, the mini-batch training / inference loops and checkpoint I/O.
 *
 * Same call surface and semantics as upstream src/auxil.c (init_network :52-297, create_dataset :299-363,
 * save/load_network :468-611, compute_error :1100-1659, train_network :1662-1972, forward_testset :1975-2039);
 * different mechanics: no per-layer device synchronisation, loss reduced on the device (one float per
 * sample comes back instead of the whole per-element loss tensor), raw gradients kept in one arena that is
 * all-reduced across GPUs (NCCL) before a fused optimizer pass.
 */
#include <math.h>
#include <string.h>
#include <time.h>
#include <sys/stat.h>
#include <sys/time.h>
#include "cianna.h"

network *networks[MAX_NETWORKS_NB];
int nb_networks = 0;
int is_init = 0;

static double now_s(void)
{
	struct timeval tv;
	gettimeofday(&tv, NULL);
	return tv.tv_sec + 1e-6 * tv.tv_usec;
}

/* ------------------------------------------------------------------ init */
void init_network(int network_number, int u_input_dim[4], int u_output_dim, float in_bias, int u_batch_size,
	const char *compute_method_string, int u_dynamic_load, const char *cuda_TC_string, int inference_only, int no_logo, int adv_size)
{
	network *net;
	const char *mode_name = "FP32C_FP32A";
	int mode = FP32C_FP32A;

	if (!is_init && !no_logo)
		printf("############################################################\n"
		       "CIANNA B200-native core (%s), API of CIANNA V-1.0.0.0\n"
		       "############################################################\n\n", cb200_version());
	if (network_number < 0 || network_number >= MAX_NETWORKS_NB) { printf("ERROR: network id out of range\n"); exit(EXIT_FAILURE); }
	if (strcmp(compute_method_string, "C_CUDA") != 0) {
		printf("ERROR: compute method %s is not available: this build only carries the CUDA (sm_100a) back-end and has no CPU fallback.\n", compute_method_string);
		exit(EXIT_FAILURE);
	}
	if (strcmp(cuda_TC_string, "off") == 0 || strcmp(cuda_TC_string, "FP32C_FP32A") == 0) { mode = FP32C_FP32A; mode_name = "FP32C_FP32A"; }
	else if (strcmp(cuda_TC_string, "on") == 0 || strcmp(cuda_TC_string, "FP16C_FP32A") == 0) { mode = FP16C_FP32A; mode_name = "FP16C_FP32A"; }
	else if (strcmp(cuda_TC_string, "BF16C_FP32A") == 0) { mode = BF16C_FP32A; mode_name = "BF16C_FP32A"; }
	else if (strcmp(cuda_TC_string, "TF32C_FP32A") == 0 || strcmp(cuda_TC_string, "FP16C_FP16A") == 0) {
		printf("ERROR: mixed precision mode %s is not provided by the B200 core (available: FP32C_FP32A, FP16C_FP32A, BF16C_FP32A).\n", cuda_TC_string);
		exit(EXIT_FAILURE);
	} else { printf("ERROR: unknown mixed_precision string %s\n", cuda_TC_string); exit(EXIT_FAILURE); }

	net = (network *)calloc(1, sizeof(network));
	networks[network_number] = net;
	net->id = network_number;
	net->compute_method = C_CUDA;
	net->dynamic_load = u_dynamic_load;
	net->use_cuda_TC = mode;
	net->dtype = mode == FP32C_FP32A ? CB200_FP32 : (mode == FP16C_FP32A ? CB200_FP16 : CB200_BF16);
	srand((unsigned)time(NULL));
	CB_CHECK(cb200_init(-1));
	if (network_number >= nb_networks) nb_networks = network_number + 1;
	is_init = 1;

	net->in_dims[0] = u_input_dim[0]; net->in_dims[1] = u_input_dim[1];
	net->in_dims[2] = u_input_dim[2]; net->in_dims[3] = u_input_dim[3];
	net->input_dim = ((size_t)u_input_dim[0]) * u_input_dim[1] * u_input_dim[2] * u_input_dim[3];
	net->output_dim = u_output_dim;
	net->input_bias = in_bias;
	if (u_batch_size > 1) { net->batch_size = u_batch_size; net->batch_param = OFF; }
	else if (u_batch_size == 1) { net->batch_size = 1; net->batch_param = SGD; printf(" Automatically switch to SGD scheme (batch_size = 1)\n"); }
	else { net->batch_size = 16; net->batch_param = FULL; printf(" Undefined batch size -> automatic value is 16\n"); }
	net->inference_only = inference_only;
	net->inference_drop_mode = AVG_MODEL;
	net->perf_eval = 1;
	net->adv_size = adv_size <= 0 ? 30 : adv_size;
	net->TC_scale_factor = 1.0f;
	net->dp_world = 1;
	/* time-seeded like upstream's generators (src/auxil.c:164, src/cuda/cuda_main.cu:1051-1052); cb_set_dropout_seed makes runs repeatable */
	net->drop_seed = (unsigned long long)time(NULL) * 0x9E3779B97F4A7C15ULL + (unsigned long long)network_number;
	net->length = net->batch_size;
	net->train_buf.localization = NO_LOC; net->test_buf.localization = NO_LOC; net->valid_buf.localization = NO_LOC;
	/* empty YOLO set-up (src/auxil.c:261-295); set_yolo_params fills it before the YOLO layer is created */
	net->y_param = (yolo_param *)calloc(1, sizeof(yolo_param));
	net->y_param->min_prior_forced_scaling = -1.0f;

	{
		size_t es = cb200_dtype_size(net->dtype);
		CB_CHECK(cb200_malloc(&net->input_raw, (size_t)net->batch_size * (net->input_dim + 1) * es));
		CB_CHECK(cb200_malloc(&net->input, (size_t)net->batch_size * net->in_dims[0] * net->in_dims[1] * net->in_dims[2] * cb200_round_channels(net->in_dims[3]) * es));
		CB_CHECK(cb200_malloc(&net->target, (size_t)net->batch_size * (net->output_dim > 0 ? net->output_dim : 1) * es));
		CB_CHECK(cb200_malloc((void **)&net->loss_dev, (size_t)net->batch_size * sizeof(float)));
		CB_CHECK(cb200_host_alloc((void **)&net->loss_host, (size_t)net->batch_size * sizeof(float)));
		CB_CHECK(cb200_malloc((void **)&net->hyper_dev, CB200_HYPER_LEN * sizeof(float)));
	}
	printf("Network (id: %d) initialized with : \nInput dimensions: %dx%dx%dx%d \nOutput dimension: %d \nBatch size: %d \n"
	       "Using CUDA (%s) compute method \nInference only: %d\n\n",
		net->id, net->in_dims[0], net->in_dims[1], net->in_dims[2], net->in_dims[3], net->output_dim, net->batch_size, mode_name, inference_only);
	if (net->dynamic_load) printf("Dynamic load ENABLED\n\n");
}

/* ------------------------------------------------------------------ datasets */
Dataset create_dataset(network *net, int nb_elem)
{
	Dataset data;
	int i, j;
	size_t es = cb200_dtype_size(net->dtype);
	size_t in_elems = (size_t)net->batch_size * (net->input_dim + 1);
	size_t out_elems = (size_t)net->batch_size * net->output_dim;
	float bias = net->input_bias;
	unsigned char bias_typed[4];

	memset(&data, 0, sizeof(data));
	data.size = nb_elem;
	data.nb_batch = (nb_elem - 1) / net->batch_size + 1;
	data.localization = HOST;
	data.input = (void **)calloc(data.nb_batch, sizeof(void *));
	data.target = (void **)calloc(data.nb_batch, sizeof(void *));
	cb200_host_cast_from_f32(bias_typed, net->dtype, &bias, 1);
	for (i = 0; i < data.nb_batch; i++) {
		CB_CHECK(cb200_host_alloc(&data.input[i], in_elems * es));
		CB_CHECK(cb200_host_alloc(&data.target[i], (out_elems ? out_elems : 1) * es));
		for (j = 0; j < net->batch_size; j++)
			memcpy((char *)data.input[i] + ((size_t)j * (net->input_dim + 1) + net->input_dim) * es, bias_typed, es);
	}
	return data;
}

static void shuffle_ws_free(Dataset *data);

void free_dataset(Dataset *data)
{
	int i;
	if (data->input == NULL) return;
	for (i = 0; i < data->nb_batch; i++) {
		cb200_host_free(data->input[i]);
		cb200_host_free(data->target[i]);
		if (data->input_device) { cb200_free(data->input_device[i]); cb200_free(data->target_device[i]); }
	}
	shuffle_ws_free(data);
	free(data->input); free(data->target);
	free(data->input_device); free(data->target_device);
	memset(data, 0, sizeof(*data));
}

void dataset_set_sample(network *net, Dataset *data, int index, const float *input, const float *target)
{
	size_t es = cb200_dtype_size(net->dtype);
	int b = index / net->batch_size, j = index % net->batch_size;
	if (index < 0 || index >= data->size) { printf("ERROR: dataset_set_sample index out of range\n"); exit(EXIT_FAILURE); }
	if (input != NULL)
		cb200_host_cast_from_f32((char *)data->input[b] + (size_t)j * (net->input_dim + 1) * es, net->dtype, input, net->input_dim);
	if (target != NULL)
		cb200_host_cast_from_f32((char *)data->target[b] + (size_t)j * net->output_dim * es, net->dtype, target, net->output_dim);
}

void dataset_upload(network *net, Dataset *data)
{
	int i;
	size_t es = cb200_dtype_size(net->dtype);
	size_t in_bytes = (size_t)net->batch_size * (net->input_dim + 1) * es;
	size_t out_bytes = (size_t)net->batch_size * net->output_dim * es;
	if (data->input_device != NULL) return;
	data->input_device = (void **)calloc(data->nb_batch, sizeof(void *));
	data->target_device = (void **)calloc(data->nb_batch, sizeof(void *));
	for (i = 0; i < data->nb_batch; i++) {
		CB_CHECK(cb200_malloc(&data->input_device[i], in_bytes));
		CB_CHECK(cb200_malloc(&data->target_device[i], out_bytes ? out_bytes : 16));
		CB_CHECK(cb200_h2d(data->input_device[i], data->input[i], in_bytes, NULL));
		CB_CHECK(cb200_h2d(data->target_device[i], data->target[i], out_bytes, NULL));
	}
	CB_CHECK(cb200_stream_sync(NULL));
	data->localization = DEVICE;
}

/* Uniform random permutation of the samples of a data set, across its batches (Fisher-Yates on whole rows, input and
 * target together).  Same effect as upstream's cuda_host_only_shuffle / cuda_host_shuffle / cuda_shuffle
 * (src/cuda/cuda_main.cu:590-768, called from src/auxil.c:1768-1789); rows are moved as bytes, so one routine serves
 * every storage type.  Device-resident sets (dynamic_load == 0) are refreshed from the shuffled host copy. */
static unsigned char *sample_row(void **batches, int batch_size, size_t row_bytes, int sample)
{
	return (unsigned char *)batches[sample / batch_size] + (size_t)(sample % batch_size) * row_bytes;
}

static void swap_rows(unsigned char *a, unsigned char *b, unsigned char *tmp, size_t bytes)
{
	memcpy(tmp, a, bytes); memcpy(a, b, bytes); memcpy(b, tmp, bytes);
}

/* ---- the same permutation applied ON THE DEVICE to a resident set (train(shuffle_gpu = 1), dynamic_load == 0):
 * upstream's cuda_shuffle (src/cuda/cuda_main.cu:642-666, src/auxil.c:1700-1709): a persistent index table takes one
 * Fisher-Yates pass on the host, goes to the device, the rows are scattered into a duplicate set and copied back.  Here
 * both moves are cb200_rows_permute launches (one warp per row); the pinned host batches are refreshed only when
 * something reads them (dataset_host_refresh). */
typedef struct {
	int nb_batch, size;
	void **dup_in, **dup_tg;                   /* host arrays of device batches */
	void *ptr_in, *ptr_tg, *ptr_dup_in, *ptr_dup_tg;   /* device arrays of nb_batch device pointers */
	int *index, *index_dev;
} shuffle_ws;

static void shuffle_ws_free(Dataset *data)
{
	shuffle_ws *ws = (shuffle_ws *)data->shuffle_ws;
	int i;
	if (ws == NULL) return;
	for (i = 0; i < ws->nb_batch; i++) { cb200_free(ws->dup_in[i]); cb200_free(ws->dup_tg[i]); }
	cb200_free(ws->ptr_in); cb200_free(ws->ptr_tg); cb200_free(ws->ptr_dup_in); cb200_free(ws->ptr_dup_tg);
	cb200_free(ws->index_dev);
	free(ws->dup_in); free(ws->dup_tg); free(ws->index); free(ws);
	data->shuffle_ws = NULL;
}

void dataset_host_refresh(network *net, Dataset *data)
{
	size_t es = cb200_dtype_size(net->dtype);
	size_t in_bytes = (size_t)net->batch_size * (net->input_dim + 1) * es, out_bytes = (size_t)net->batch_size * net->output_dim * es;
	int i;
	if (!data->host_stale || data->input_device == NULL) return;
	for (i = 0; i < data->nb_batch; i++) {
		CB_CHECK(cb200_d2h(data->input[i], data->input_device[i], in_bytes, NULL));
		if (out_bytes) CB_CHECK(cb200_d2h(data->target[i], data->target_device[i], out_bytes, NULL));
	}
	CB_CHECK(cb200_stream_sync(NULL));
	data->host_stale = 0;
}

void shuffle_dataset_device(network *net, Dataset *data)
{
	size_t es = cb200_dtype_size(net->dtype);
	size_t in_row = (net->input_dim + 1) * es, out_row = (size_t)net->output_dim * es;
	shuffle_ws *ws = (shuffle_ws *)data->shuffle_ws;
	int i;
	if (data->input_device == NULL) { printf("\nERROR: shuffle_dataset_device: the data set has no device-resident copy\n"); exit(EXIT_FAILURE); }
	if (ws == NULL) {
		size_t pb = (size_t)data->nb_batch * sizeof(void *);
		ws = (shuffle_ws *)calloc(1, sizeof(shuffle_ws));
		ws->nb_batch = data->nb_batch; ws->size = data->size;
		ws->dup_in = (void **)calloc(data->nb_batch, sizeof(void *));
		ws->dup_tg = (void **)calloc(data->nb_batch, sizeof(void *));
		for (i = 0; i < data->nb_batch; i++) {
			CB_CHECK(cb200_malloc(&ws->dup_in[i], (size_t)net->batch_size * in_row));
			CB_CHECK(cb200_malloc(&ws->dup_tg[i], out_row ? (size_t)net->batch_size * out_row : 16));
		}
		CB_CHECK(cb200_malloc(&ws->ptr_in, pb)); CB_CHECK(cb200_malloc(&ws->ptr_tg, pb));
		CB_CHECK(cb200_malloc(&ws->ptr_dup_in, pb)); CB_CHECK(cb200_malloc(&ws->ptr_dup_tg, pb));
		CB_CHECK(cb200_h2d(ws->ptr_in, data->input_device, pb, NULL)); CB_CHECK(cb200_h2d(ws->ptr_tg, data->target_device, pb, NULL));
		CB_CHECK(cb200_h2d(ws->ptr_dup_in, ws->dup_in, pb, NULL)); CB_CHECK(cb200_h2d(ws->ptr_dup_tg, ws->dup_tg, pb, NULL));
		ws->index = (int *)malloc((size_t)data->size * sizeof(int));
		for (i = 0; i < data->size; i++) ws->index[i] = i;
		CB_CHECK(cb200_malloc((void **)&ws->index_dev, (size_t)data->size * sizeof(int)));
		CB_CHECK(cb200_stream_sync(NULL));
		data->shuffle_ws = ws;
	}
	for (i = 0; i < data->size - 1; i++) {
		int j = i + (int)((rand() / ((double)RAND_MAX + 1.0)) * (double)(data->size - i));
		int t = ws->index[i]; ws->index[i] = ws->index[j]; ws->index[j] = t;
	}
	CB_CHECK(cb200_h2d(ws->index_dev, ws->index, (size_t)data->size * sizeof(int), NULL));
	CB_CHECK(cb200_rows_permute((void *const *)ws->ptr_dup_in, (void *const *)ws->ptr_in, ws->index_dev, data->size, net->batch_size, in_row, NULL));
	CB_CHECK(cb200_rows_permute((void *const *)ws->ptr_in, (void *const *)ws->ptr_dup_in, NULL, data->size, net->batch_size, in_row, NULL));
	if (out_row) {
		CB_CHECK(cb200_rows_permute((void *const *)ws->ptr_dup_tg, (void *const *)ws->ptr_tg, ws->index_dev, data->size, net->batch_size, out_row, NULL));
		CB_CHECK(cb200_rows_permute((void *const *)ws->ptr_tg, (void *const *)ws->ptr_dup_tg, NULL, data->size, net->batch_size, out_row, NULL));
	}
	CB_CHECK(cb200_stream_sync(NULL));      /* ws->index is rewritten by the next call */
	data->host_stale = 1;
}

void shuffle_dataset(network *net, Dataset *data)
{
	size_t es = cb200_dtype_size(net->dtype);
	size_t in_row = (net->input_dim + 1) * es, out_row = (size_t)net->output_dim * es;
	unsigned char *tmp = (unsigned char *)malloc(in_row > out_row ? in_row : out_row);
	int i;
	dataset_host_refresh(net, data);
	for (i = 0; i < data->size - 1; i++) {
		int j = i + (int)((rand() / ((double)RAND_MAX + 1.0)) * (double)(data->size - i));
		if (j == i) continue;
		swap_rows(sample_row(data->input, net->batch_size, in_row, i), sample_row(data->input, net->batch_size, in_row, j), tmp, in_row);
		if (out_row)
			swap_rows(sample_row(data->target, net->batch_size, out_row, i), sample_row(data->target, net->batch_size, out_row, j), tmp, out_row);
	}
	free(tmp);
	if (data->input_device != NULL) {
		for (i = 0; i < data->nb_batch; i++) {
			CB_CHECK(cb200_h2d(data->input_device[i], data->input[i], (size_t)net->batch_size * in_row, NULL));
			if (out_row) CB_CHECK(cb200_h2d(data->target_device[i], data->target[i], (size_t)net->batch_size * out_row, NULL));
		}
		CB_CHECK(cb200_stream_sync(NULL));
	}
}

void cb_net_io_dims(network *net, long long *out3) { out3[0] = (long long)net->input_dim; out3[1] = net->output_dim; out3[2] = net->dtype; }

/* raw bytes (storage type of the network) of one sample's input (which = 0, input_dim + 1 values) or target row
 * (which = 1); from_device != 0 reads the device-resident copy instead of the host one (tests) */
void cb_dataset_read_row(network *net, Dataset *data, int index, int which, int from_device, void *dst)
{
	size_t es = cb200_dtype_size(net->dtype);
	size_t row = which == 0 ? (net->input_dim + 1) * es : (size_t)net->output_dim * es;
	int b = index / net->batch_size, j = index % net->batch_size;
	if (index < 0 || index >= data->size) { printf("ERROR: cb_dataset_read_row index out of range\n"); exit(EXIT_FAILURE); }
	if (from_device) {
		if (data->input_device == NULL) { printf("ERROR: the data set has no device-resident copy\n"); exit(EXIT_FAILURE); }
		CB_CHECK(cb200_d2h(dst, (char *)(which == 0 ? data->input_device[b] : data->target_device[b]) + (size_t)j * row, row, NULL));
		CB_CHECK(cb200_stream_sync(NULL));
	} else {
		dataset_host_refresh(net, data);
		memcpy(dst, (char *)(which == 0 ? data->input[b] : data->target[b]) + (size_t)j * row, row);
	}
}

/* ------------------------------------------------------------------ training preparation */
int cb_dp_plan(int n, const size_t *len, size_t head, size_t *offset, size_t *begin, size_t *blen, int *first)
{
	size_t total = (head + 63) & ~(size_t)63, weights, acc, target;
	int i, nb = 0, hi;
	const size_t tail_cap = (size_t)1 << 20;      /* the last bucket to complete sits on the critical path: <= 4 MB */
	for (i = 0; i < n; i++) { offset[i] = total; total += (len[i] + 63) & ~(size_t)63; }
	if (n == 0) { if (head > 0) { begin[0] = 0; blen[0] = total; first[0] = -1; return 1; } return 0; }
	weights = total - offset[0];
	target = weights / 4 > tail_cap ? weights / 4 : tail_cap;      /* ~4 large buckets: launch latency, not link count */
	/* walk from the last slice down (the order the backward sweep produces them) */
	hi = n;             /* current bucket = slices [i, hi) */
	acc = 0;
	for (i = n - 1; i >= 0; i--) {
		size_t below = offset[i] - offset[0];      /* floats of the slices under i */
		acc += (len[i] + 63) & ~(size_t)63;
		if (i == 0) break;
		/* close when large enough, or when what remains below is the small tail and this bucket is not tiny itself */
		if ((acc >= target || (below <= tail_cap && acc > tail_cap)) && nb < CB_DP_MAX_BUCKETS - 1) {
			begin[nb] = offset[i]; blen[nb] = acc; first[nb] = i; nb++;
			hi = i; acc = 0;
		}
	}
	(void)hi;
	/* the final bucket also carries the group-norm sums at the head of the arena */
	begin[nb] = 0; blen[nb] = offset[0] + acc; first[nb] = 0; nb++;
	return nb;
}

static void prepare_training(network *net)
{
	int k, n = 0, b;
	size_t head = 0, total;
	size_t len[MAX_LAYERS_NB], offset[MAX_LAYERS_NB];
	int owner[MAX_LAYERS_NB], first[CB_DP_MAX_BUCKETS];
	if (net->training_ready) return;
	if (net->inference_only) { printf("\nERROR: network was created in inference only mode, it cannot be trained.\n"); exit(EXIT_FAILURE); }
	/* the (tiny) group-norm gradient sums sit together at the head of the arena and travel with the last bucket */
	for (k = 0; k < net->nb_layers; k++) {
		layer *l = net->net_layers[k];
		if (l->type == NORM) {
			norm_param *p = (norm_param *)l->param;
			p->grad_offset = head;
			head += 2 * (size_t)p->nb_group;
		} else if (l->type == CONV) {
			conv_param *p = (conv_param *)l->param;
			p->grad_len = cb200_conv_grad_elems(&p->desc) + p->desc.out_c;
			len[n] = p->grad_len; owner[n++] = k;
		} else if (l->type == DENSE) {
			dense_param *p = (dense_param *)l->param;
			p->grad_len = cb200_conv_grad_elems(&p->desc) + p->desc.out_c;
			len[n] = p->grad_len; owner[n++] = k;
		}
	}
	net->dp_nb_bucket = cb_dp_plan(n, len, head, offset, net->dp_bucket_begin, net->dp_bucket_len, first);
	total = n > 0 ? offset[n - 1] + ((len[n - 1] + 63) & ~(size_t)63) : ((head + 63) & ~(size_t)63);
	for (b = 0; b < net->dp_nb_bucket; b++) {
		net->dp_bucket_trigger[b] = first[b] >= 0 ? owner[first[b]] : -1;
	}
	net->grad_arena_len = total;
	CB_CHECK(cb200_malloc((void **)&net->grad_arena, (total ? total : 1) * sizeof(float)));
	CB_CHECK(cb200_memset(net->grad_arena, 0, (total ? total : 1) * sizeof(float), NULL));
	n = 0;
	for (k = 0; k < net->nb_layers; k++) {
		layer *l = net->net_layers[k];
		if (l->type == CONV) {
			conv_param *p = (conv_param *)l->param;
			p->grad_offset = offset[n++];
			p->w.grad = net->grad_arena + p->grad_offset;
			p->w.grad_b = p->w.grad + cb200_conv_grad_elems(&p->desc);
		} else if (l->type == DENSE) {
			dense_param *p = (dense_param *)l->param;
			p->grad_offset = offset[n++];
			p->w.grad = net->grad_arena + p->grad_offset;
			p->w.grad_b = p->w.grad + cb200_conv_grad_elems(&p->desc);
		} else if (l->type == NORM) {
			norm_param *p = (norm_param *)l->param;
			p->gsum = net->grad_arena + p->grad_offset;
		}
	}
	{
		const char *e = getenv("CB200_NO_WGRAD_STREAM");
		if (!(e != NULL && e[0] != '\0' && e[0] != '0')) CB_CHECK(cb200_stream_create_low_priority(&net->wgrad_stream));
	}
	net->training_ready = 1;
}

/* data parallel: layer `current` has just enqueued its weight gradient (or skipped it: frozen).  If it is the lowest
 * layer of an exchange bucket, everything the bucket holds is now enqueued - convolution weight gradients on the
 * weight-gradient stream, dense ones, bias columns written by a following norm layer and the group-norm sums on the
 * compute stream - and the whole slice goes out in ONE all-reduce on the communication stream. */
void cb_dp_layer_done(network *net, layer *current)
{
	int b;
	if (net->dp_world <= 1) return;
	for (b = 0; b < net->dp_nb_bucket; b++) {
		if (net->dp_bucket_trigger[b] != current->index) continue;
		if (net->wgrad_stream != NULL) CB_CHECK(cb200_dp_after(NULL));
		CB_CHECK(cb200_dp_allreduce(net->grad_arena + net->dp_bucket_begin[b], net->dp_bucket_len[b], net->wgrad_stream));
	}
}

static void last_layer_dims(network *net, int *c, int *h, int *w)
{
	layer *last = net->net_layers[net->nb_layers - 1];
	*c = last->out_c; *h = last->out_h; *w = last->out_w;
}

static void set_hyper(network *net, float lr, float momentum, float weight_decay)
{
	float h[CB200_HYPER_LEN];
	memset(h, 0, sizeof(h));
	net->learning_rate = lr; net->momentum = momentum; net->weight_decay = weight_decay;
	h[0] = lr / (float)(net->batch_size * net->dp_world);
	h[1] = momentum;
	h[2] = lr * weight_decay;
	h[3] = net->TC_scale_factor;
	h[4] = lr;
	CB_CHECK(cb200_h2d(net->hyper_dev, h, sizeof(h), NULL));
}

/* ------------------------------------------------------------------ one mini-batch */
void cb_load_batch_typed(network *net, const void *input_typed, const void *target_typed)
{
	size_t es = cb200_dtype_size(net->dtype);
	CB_CHECK(cb200_h2d(net->input_raw, input_typed, (size_t)net->batch_size * (net->input_dim + 1) * es, NULL));
	if (target_typed != NULL && net->output_dim > 0)
		CB_CHECK(cb200_h2d(net->target, target_typed, (size_t)net->batch_size * net->output_dim * es, NULL));
}

void cb_load_batch(network *net, const float *input, const float *target)
{
	size_t es = cb200_dtype_size(net->dtype);
	size_t n_in = (size_t)net->batch_size * (net->input_dim + 1), n_out = (size_t)net->batch_size * net->output_dim;
	void *ti = malloc(n_in * es), *tt = malloc((n_out ? n_out : 1) * es);
	cb200_host_cast_from_f32(ti, net->dtype, input, n_in);
	if (target != NULL) cb200_host_cast_from_f32(tt, net->dtype, target, n_out);
	cb_load_batch_typed(net, ti, target != NULL ? tt : NULL);
	CB_CHECK(cb200_stream_sync(NULL));
	free(ti); free(tt);
}

static void use_device_batch(network *net, const void *input_dev)
{
	/* dataset layout -> channels-last (or -> first-layer patch rows) */
	if (net->patch_desc != NULL && net->patch_desc->input_is_patches == 2) {
		net->input = (void *)input_dev;      /* the first layer reads the dataset batch directly */
	} else if (net->patch_desc != NULL) {
		const cb200_conv_desc *d = net->patch_desc;
		CB_CHECK(cb200_import_input_patches(net->input, input_dev, net->dtype, net->batch_size, d->in_c, d->in_h, d->in_w,
			d->f_h, d->f_w, d->stride_h, d->stride_w, d->pad_h, d->pad_w, d->out_h, d->out_w, d->bias_value, NULL));
	} else
		CB_CHECK(cb200_import_input(net->input, input_dev, net->dtype, net->batch_size, net->in_dims[3], net->in_dims[1] * net->in_dims[2], net->in_dims[0], NULL));
}

/* ---- per-layer timing sample (perf_eval): slot s of pass p (0 forward, 1 backward, 2 optimizer) */
static void perf_mark(network *net, int pass, int slot)
{
	if (net->perf_sample) CB_CHECK(cb200_event_record(net->perf_ev[pass * (net->nb_layers + 1) + slot], NULL));
}

static void perf_begin_sample(network *net)
{
	int i, n = 3 * (net->nb_layers + 1);
	if (net->perf_ev == NULL) {
		net->perf_ev = (void **)calloc(n, sizeof(void *));
		for (i = 0; i < n; i++) CB_CHECK(cb200_event_create(&net->perf_ev[i]));
		net->fwd_perf = (double *)calloc(net->nb_layers, sizeof(double));
		net->back_perf = (double *)calloc(net->nb_layers, sizeof(double));
	}
	net->perf_sample = 1;
}

/* after the sampled batch has completed: forward slot k .. k+1 brackets layer k, backward slot i .. i+1 brackets layer
 * nb_layers-1-i (slot 0 .. 1 also holds the output error, like upstream), optimizer slot k .. k+1 layer k's update */
static void perf_end_sample(network *net)
{
	int k, L = net->nb_layers;
	float ms;
	net->perf_sample = 0;
	for (k = 0; k < L; k++) {
		CB_CHECK(cb200_event_elapsed_ms(net->perf_ev[k], net->perf_ev[k + 1], &ms));
		net->fwd_perf[k] += 1e3 * ms;
		CB_CHECK(cb200_event_elapsed_ms(net->perf_ev[(L + 1) + k], net->perf_ev[(L + 1) + k + 1], &ms));
		net->back_perf[L - 1 - k] += 1e3 * ms;
		CB_CHECK(cb200_event_elapsed_ms(net->perf_ev[2 * (L + 1) + k], net->perf_ev[2 * (L + 1) + k + 1], &ms));
		net->back_perf[k] += 1e3 * ms;
	}
	net->perf_n++;
}

void cb_forward(network *net, int length, int is_inference)
{
	int k;
	net->length = length;
	net->is_inference = is_inference;
	use_device_batch(net, net->input_raw);
	for (k = 0; k < net->nb_layers; k++)
		net->net_layers[k]->forward(net->net_layers[k]);
}

static void output_deriv_error(network *net, const void *target_dev)
{
	layer *last = net->net_layers[net->nb_layers - 1];
	if (last->activation_type == YOLO) {
		/* target association + error signal in one pass (src/cuda/cuda_activ_functions.cu:2369-2381) */
		yolo_param *y = (yolo_param *)last->activ_param;
		y->desc.length = net->length;
		CB_CHECK(cb200_yolo_delta(&y->desc, last->delta_o, last->output, target_dev, net->TC_scale_factor,
			(long long)net->iter * net->train.size, y->seed, y->step++, y->box_state_dev, y->workspace, NULL));
		return;
	}
	/* quadratic (LIN / RELU / LOGI outputs) and cross-entropy (SMAX) share delta = (o - t) * S upstream */
	/* ... and RELU / LOGI outputs then take their own derivative (src/cuda/cuda_activ_functions.cu:2221-2233,2273-2285) */
	CB_CHECK(cb200_output_delta_activ(last->delta_o, last->output, target_dev, net->dtype, net->batch_size, net->length,
		last->out_c, last->out_h, last->out_w, net->TC_scale_factor, &last->activ, NULL));
}

static void output_error(network *net, const void *target_dev)
{
	layer *last = net->net_layers[net->nb_layers - 1];
	if (last->activation_type == YOLO) {
		yolo_param *y = (yolo_param *)last->activ_param;
		y->desc.length = net->length;
		CB_CHECK(cb200_yolo_loss(&y->desc, net->loss_dev, y->parts_dev, y->monitor_dev, last->output, target_dev, y->workspace, NULL));
		return;
	}
	CB_CHECK(cb200_output_loss(net->loss_dev, last->output, target_dev, net->dtype, net->batch_size, net->length,
		last->out_c, last->out_h, last->out_w, last->activation_type == SOFTMAX ? 1 : 0, NULL));
}

static int update_plan_on = -1;   /* -1: take CB200_UPDATE_PLAN (default on) at first use */

void cb_set_update_plan(int on) { update_plan_on = on ? 1 : 0; }

static int update_plan_enabled(void)
{
	if (update_plan_on < 0) { const char *e = getenv("CB200_UPDATE_PLAN"); update_plan_on = (e != NULL && e[0] == '0') ? 0 : 1; }
	return update_plan_on;
}

/* (re)build the plan when the set of layers it covers may have changed: first use, a layer frozen / unfrozen, the
 * data-parallel world joined since */
static void update_plan_refresh(network *net, int skip_first)
{
	const cb200_conv_desc *descs[MAX_LAYERS_NB];
	const cb200_conv_weights *ws[MAX_LAYERS_NB];
	cb200_norm_update_ref norms[MAX_LAYERS_NB];
	unsigned sig = 2166136261u;
	int k, nc = 0, nn = 0;
	const int dp = cb200_dp_world() > 1;
	for (k = 0; k < net->nb_layers; k++) {
		layer *l = net->net_layers[k];
		sig = (sig ^ (unsigned)(l->frozen ? 2 : 1)) * 16777619u;
		if (l->type == CONV) sig = (sig ^ (unsigned)(size_t)((conv_param *)l->param)->w.master) * 16777619u;
		else if (l->type == NORM) sig = (sig ^ (unsigned)(size_t)((norm_param *)l->param)->gamma) * 16777619u;
	}
	sig = (sig ^ (unsigned)dp) * 16777619u;
	sig = (sig ^ (unsigned)(skip_first ? 7 : 3)) * 16777619u;
	sig = (sig ^ (unsigned)(size_t)net->grad_arena) * 16777619u;
	if (sig == 0) sig = 1;
	if (net->update_plan_sig == sig) return;
	if (net->update_plan != NULL) { cb200_update_plan_destroy(net->update_plan); net->update_plan = NULL; }
	memset(net->in_update_plan, 0, sizeof(net->in_update_plan));
	for (k = 0; k < net->nb_layers; k++) {
		layer *l = net->net_layers[k];
		if (l->frozen || (skip_first && k == 0)) continue;
		if (l->type == CONV) {
			conv_param *p = (conv_param *)l->param;
			if (!cb200_update_plan_accepts(&p->desc)) continue;
			descs[nc] = &p->desc; ws[nc] = &p->w; nc++;
			net->in_update_plan[k] = 1;
		} else if (l->type == NORM) {
			norm_param *p = (norm_param *)l->param;
			norms[nn].d_gamma = p->d_gamma; norms[nn].d_beta = p->d_beta; norms[nn].gsum = p->gsum;
			norms[nn].gamma = p->gamma; norms[nn].beta = p->beta; norms[nn].gamma_upd = p->gamma_update; norms[nn].beta_upd = p->beta_update;
			norms[nn].batch = p->desc.batch; norms[nn].nb_group = p->desc.nb_group; norms[nn].set_off = p->desc.set_off;
			norms[nn].reduce = dp ? 0 : 1;
			nn++;
			net->in_update_plan[k] = 1;
		}
	}
	if (nc + nn >= 2) CB_CHECK(cb200_update_plan_create(&net->update_plan, net->dtype, descs, ws, nc, norms, nn));
	else memset(net->in_update_plan, 0, sizeof(net->in_update_plan));
	net->update_plan_sig = sig;
}

static void update_one_layer(network *net, layer *l)
{
	if (l->type == CONV) {
		conv_param *p = (conv_param *)l->param;
		CB_CHECK(cb200_conv_update(&p->desc, &p->w, net->hyper_dev, 0, NULL));
	} else if (l->type == DENSE) {
		dense_param *p = (dense_param *)l->param;
		CB_CHECK(cb200_dense_update(&p->desc, &p->w, net->hyper_dev, NULL));
	} else if (l->type == NORM) {
		norm_param *p = (norm_param *)l->param;
		if (cb200_dp_world() > 1)
			CB_CHECK(cb200_norm_update(&p->desc, p->gamma, p->beta, p->gamma_update, p->beta_update, p->gsum, net->hyper_dev, NULL));
		else
			CB_CHECK(cb200_norm_reduce_update(&p->desc, p->d_gamma, p->d_beta, p->gsum, p->gamma, p->beta, p->gamma_update,
				p->beta_update, net->hyper_dev, NULL));
	}
}

static int early_updates_enabled(void)
{
	static int on = -1;
	if (on < 0) { const char *e = getenv("CB200_EARLY_UPDATES"); on = (e != NULL && e[0] == '0') ? 0 : 1; }
	return on;
}

/* backward_pass calls this right before the FIRST layer's backprop: from here on the side stream only receives that
 * layer's weight gradient */
static void mark_upper_wgrads(network *net)
{
	layer *first = net->net_layers[0];
	net->upper_marked = 0;
	if (net->wgrad_stream == NULL || net->dp_world > 1 || net->perf_sample || net->nb_layers < 2 || !early_updates_enabled()) return;
	if (first->type != CONV || first->frozen) return;
	if (net->upper_ev == NULL) CB_CHECK(cb200_event_create(&net->upper_ev));
	CB_CHECK(cb200_event_record(net->upper_ev, net->wgrad_stream));
	net->upper_marked = 1;
}

static void apply_updates(network *net)
{
	int k, planned = 0;
	if (net->upper_marked) {
		/* every layer above the first: their gradients are complete at `upper_ev` (side stream) / in stream order (compute
		 * stream); these launches overlap the first layer's weight gradient, which is joined afterwards */
		net->upper_marked = 0;
		CB_CHECK(cb200_stream_wait_event(NULL, net->upper_ev));
		perf_mark(net, 2, 0);
		if (update_plan_enabled()) {
			update_plan_refresh(net, 1);
			if (net->update_plan != NULL) { CB_CHECK(cb200_update_plan_run(net->update_plan, net->hyper_dev, NULL)); planned = 1; }
		}
		for (k = 1; k < net->nb_layers; k++) {
			layer *l = net->net_layers[k];
			if (l->frozen || (planned && net->in_update_plan[k])) continue;
			update_one_layer(net, l);
		}
		CB_CHECK(cb200_stream_wait(NULL, net->wgrad_stream));      /* the first layer's weight gradient */
		update_one_layer(net, net->net_layers[0]);
		perf_mark(net, 2, net->nb_layers);
		return;
	}
	if (net->wgrad_stream != NULL) CB_CHECK(cb200_stream_wait(NULL, net->wgrad_stream));   /* all weight gradients are in */
	if (net->dp_world > 1) {
		/* a network without conv / dense layers below its norm layers has no trigger layer: exchange the head now */
		if (net->dp_nb_bucket > 0 && net->dp_bucket_trigger[net->dp_nb_bucket - 1] < 0)
			CB_CHECK(cb200_dp_allreduce(net->grad_arena, net->dp_bucket_len[net->dp_nb_bucket - 1], NULL));
		CB_CHECK(cb200_dp_join(NULL));
	}
	perf_mark(net, 2, 0);
	/* every eligible conv layer and every group-norm layer in three launches (update_plan.cu); a perf_eval sample keeps
	 * the layer-by-layer calls so that every kernel lands between its own layer's events */
	if (!net->perf_sample && update_plan_enabled()) {
		update_plan_refresh(net, 0);
		if (net->update_plan != NULL) {
			CB_CHECK(cb200_update_plan_run(net->update_plan, net->hyper_dev, NULL));
			planned = 1;
		}
	}
	for (k = 0; k < net->nb_layers; k++) {
		layer *l = net->net_layers[k];
		if (k > 0) perf_mark(net, 2, k);
		if (l->frozen || (planned && net->in_update_plan[k])) continue;
		if (l->type == CONV) {
			conv_param *p = (conv_param *)l->param;
			CB_CHECK(cb200_conv_update(&p->desc, &p->w, net->hyper_dev, 0, NULL));
		} else if (l->type == DENSE) {
			dense_param *p = (dense_param *)l->param;
			CB_CHECK(cb200_dense_update(&p->desc, &p->w, net->hyper_dev, NULL));
		} else if (l->type == NORM) {
			norm_param *p = (norm_param *)l->param;
			if (cb200_dp_world() > 1)
				CB_CHECK(cb200_norm_update(&p->desc, p->gamma, p->beta, p->gamma_update, p->beta_update, p->gsum, net->hyper_dev, NULL));
			else
				CB_CHECK(cb200_norm_reduce_update(&p->desc, p->d_gamma, p->d_beta, p->gsum, p->gamma, p->beta, p->gamma_update,
					p->beta_update, net->hyper_dev, NULL));
		}
	}
	perf_mark(net, 2, net->nb_layers);
}

static void backward_pass(network *net, const void *target_dev)
{
	int k;
	perf_mark(net, 1, 0);
	output_deriv_error(net, target_dev);
	for (k = net->nb_layers - 1; k >= 0; k--) {
		if (k == 0) mark_upper_wgrads(net);
		net->net_layers[k]->backprop(net->net_layers[k]);
		perf_mark(net, 1, net->nb_layers - k);
	}
	apply_updates(net);
}

/* the pieces of one step under their own names, for the upstream-side back-end shim (bridge.c), which is driven layer
 * by layer from upstream's own loops */
void cb_use_device_batch(network *net, const void *input_dev) { use_device_batch(net, input_dev); }
void cb_prepare_training(network *net) { prepare_training(net); }
void cb_set_hyper(network *net, float lr, float momentum, float weight_decay) { set_hyper(net, lr, momentum, weight_decay); }
void cb_apply_updates(network *net) { apply_updates(net); }
void cb_output_deriv_error(network *net, const void *target_dev) { output_deriv_error(net, target_dev); }
void cb_output_error(network *net, const void *target_dev) { output_error(net, target_dev); }

void cb_backward(network *net, float lr, float momentum, float weight_decay)
{
	prepare_training(net);
	set_hyper(net, lr, momentum, weight_decay);
	backward_pass(net, net->target);
}

float cb_batch_loss(network *net)
{
	int k;
	double s = 0.0;
	output_error(net, net->target);
	CB_CHECK(cb200_d2h(net->loss_host, net->loss_dev, (size_t)net->batch_size * sizeof(float), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
	for (k = 0; k < net->length; k++) s += net->loss_host[k];
	return net->length > 0 ? (float)(s / net->length) : 0.0f;
}

void cb_train_step(network *net, float lr, float momentum, float weight_decay)
{
	cb_forward(net, net->batch_size, 0);
	cb_backward(net, lr, momentum, weight_decay);
}

void cb_sync(void) { CB_CHECK(cb200_device_sync()); }

/* mean per-sample loss of the LAST training step that was enqueued (train_one_batch's monitor copy), after a sync */
float cb_last_step_loss(network *net)
{
	int k;
	double s = 0.0;
	CB_CHECK(cb200_stream_sync(NULL));
	for (k = 0; k < net->length; k++) s += net->loss_host[k];
	return net->length > 0 ? (float)(s / net->length) : 0.0f;
}

void cb_dp_unique_id(void *id128) { CB_CHECK(cb200_dp_unique_id(id128)); }
extern void dense_refresh_operands(layer *cur);

/* Every replica starts from rank 0's parameters and optimizer state: the initialisers are time-seeded per process
 * (init_network) and a rank may have loaded another checkpoint, and nothing but gradients is exchanged afterwards.
 * Broadcast the FP32 master weights and momentum buffers (conv / dense) and gamma / beta with their update buffers
 * (norm), then rebuild the 16-bit operand copies from the masters.  Call after the layers exist (and again after a
 * load) - replicas that diverged before this call are made identical, not detected. */
void cb_dp_sync_parameters(network *net)
{
	int k;
	if (net->dp_world <= 1) return;
	for (k = 0; k < net->nb_layers; k++) {
		layer *l = net->net_layers[k];
		if (l->type == CONV) {
			conv_param *p = (conv_param *)l->param;
			size_t bytes = cb200_conv_master_elems(&p->desc) * sizeof(float);
			CB_CHECK(cb200_dp_broadcast(p->w.master, bytes, 0, NULL));
			if (p->w.moment != NULL) CB_CHECK(cb200_dp_broadcast(p->w.moment, bytes, 0, NULL));
			CB_CHECK(cb200_conv_prepare_weights(&p->desc, &p->w, NULL));
		} else if (l->type == DENSE) {
			dense_param *p = (dense_param *)l->param;
			size_t bytes = (size_t)p->in_size * (p->nb_neurons + 1) * sizeof(float);
			CB_CHECK(cb200_dp_broadcast(p->w.master, bytes, 0, NULL));
			if (p->w.moment != NULL) CB_CHECK(cb200_dp_broadcast(p->w.moment, bytes, 0, NULL));
			dense_refresh_operands(l);
		} else if (l->type == NORM) {
			norm_param *p = (norm_param *)l->param;
			size_t bytes = (size_t)p->nb_group * sizeof(float);
			CB_CHECK(cb200_dp_broadcast(p->gamma, bytes, 0, NULL));
			CB_CHECK(cb200_dp_broadcast(p->beta, bytes, 0, NULL));
			if (p->gamma_update != NULL) CB_CHECK(cb200_dp_broadcast(p->gamma_update, bytes, 0, NULL));
			if (p->beta_update != NULL) CB_CHECK(cb200_dp_broadcast(p->beta_update, bytes, 0, NULL));
		}
	}
	CB_CHECK(cb200_stream_sync(NULL));
}

void cb_dp_init(network *net, const void *id128, int rank, int world)
{
	CB_CHECK(cb200_dp_init(id128, rank, world));
	net->dp_world = world;
	net->drop_seed += 0xD1B54A32D192ED03ULL * (unsigned long long)(rank + 1);   /* each rank draws its own dropout masks */
	cb_dp_sync_parameters(net);
}

/* ------------------------------------------------------------------ training loop */
/* ---- dynamic_load staging: batch j+1 travels host -> device on a copy stream while step j computes.
 * Replaces the blocking cudaMemcpy per batch of upstream (src/auxil.c:1811-1819). */
static void stage_init(network *net)
{
	size_t es = cb200_dtype_size(net->dtype);
	int s;
	if (net->copy_stream != NULL) return;
	CB_CHECK(cb200_stream_create(&net->copy_stream));
	for (s = 0; s < 2; s++) {
		CB_CHECK(cb200_malloc(&net->stage_in[s], (size_t)net->batch_size * (net->input_dim + 1) * es));
		CB_CHECK(cb200_malloc(&net->stage_tg[s], (size_t)net->batch_size * (net->output_dim > 0 ? net->output_dim : 1) * es));
		net->staged_src[s] = NULL;
	}
	CB_CHECK(cb200_stream_sync(NULL));
}

static void stage_issue(network *net, int slot, const void *in_host, const void *tg_host)
{
	size_t es = cb200_dtype_size(net->dtype);
	CB_CHECK(cb200_h2d(net->stage_in[slot], in_host, (size_t)net->batch_size * (net->input_dim + 1) * es, net->copy_stream));
	if (tg_host != NULL && net->output_dim > 0)
		CB_CHECK(cb200_h2d(net->stage_tg[slot], tg_host, (size_t)net->batch_size * net->output_dim * es, net->copy_stream));
	net->staged_src[slot] = in_host;
}

static void stage_invalidate(network *net) { net->staged_src[0] = NULL; net->staged_src[1] = NULL; }

/* returns the slot holding batch (in_host, tg_host), copying it now if it was not prefetched, then starts the copy of
 * the next batch into the other slot; on return the compute stream is ordered after the current batch's copy */
static int stage_acquire(network *net, const void *in_host, const void *tg_host, const void *next_in, const void *next_tg)
{
	int slot;
	stage_init(net);
	slot = net->stage_slot;
	if (net->staged_src[slot] != in_host) {
		/* not prefetched (first step, or the caller jumped): the slot may still be read by earlier compute work */
		CB_CHECK(cb200_stream_wait(net->copy_stream, NULL));
		stage_issue(net, slot, in_host, tg_host);
	}
	CB_CHECK(cb200_stream_wait(NULL, net->copy_stream));        /* compute waits for this batch */
	if (next_in != NULL) {
		CB_CHECK(cb200_stream_wait(net->copy_stream, NULL));    /* the other slot's last reader (previous step) is queued */
		stage_issue(net, slot ^ 1, next_in, next_tg);
	}
	net->stage_slot = slot ^ 1;
	return slot;
}

/* One mini-batch of the training loop (body of upstream's batch loop, src/auxil.c:1797-1917): host->device copy of
 * the batch when it is not device-resident, layout import, forward sweep, loss monitor (per-sample sums, async
 * device->host), backward sweep, gradient exchange and optimizer.  Everything is enqueued; the caller decides
 * when to synchronise. */
static void train_one_batch(network *net, Dataset *data, int j, int resident, int j_next)
{
	int k;
	const void *tgt;
	net->is_inference = 0;
	net->length = (j == data->nb_batch - 1 && data->size % net->batch_size > 0) ? data->size % net->batch_size : net->batch_size;
	if (!resident) {
		int slot = stage_acquire(net, data->input[j], data->target[j],
			j_next >= 0 ? data->input[j_next] : NULL, j_next >= 0 ? data->target[j_next] : NULL);
		use_device_batch(net, net->stage_in[slot]);
		tgt = net->stage_tg[slot];
	} else {
		use_device_batch(net, data->input_device[j]);
		tgt = data->target_device[j];
	}
	perf_mark(net, 0, 0);
	for (k = 0; k < net->nb_layers; k++) {
		net->net_layers[k]->forward(net->net_layers[k]);
		perf_mark(net, 0, k + 1);
	}
	/* the loss monitor reads the forward output, so it can be queued before the backward sweep */
	output_error(net, tgt);
	CB_CHECK(cb200_d2h(net->loss_host, net->loss_dev, (size_t)net->batch_size * sizeof(float), NULL));
	backward_pass(net, tgt);
}

/* exactly `nsteps` training steps cycling over the TRAIN dataset's batches with the current hyper-parameters
 * (benchmark / profiling entry: the same per-batch work as train_network, without the printing) */
void cb_train_steps(network *net, int nsteps, float lr, float momentum, float weight_decay, int resident, int sync_each_step)
{
	int s;
	if (net->train.input == NULL) { printf("\nERROR: no TRAIN dataset defined\n"); exit(EXIT_FAILURE); }
	prepare_training(net);
	if (resident) dataset_upload(net, &net->train);
	stage_invalidate(net);      /* host batches may have been rewritten since the last call */
	set_hyper(net, lr, momentum, weight_decay);
	for (s = 0; s < nsteps; s++) {
		train_one_batch(net, &net->train, s % net->train.nb_batch, resident, s + 1 < nsteps ? (s + 1) % net->train.nb_batch : -1);
		if (sync_each_step) CB_CHECK(cb200_stream_sync(NULL));
	}
}

/* `nsteps` forward-only passes over the TEST dataset's batches (inference throughput) */
void cb_forward_steps(network *net, int nsteps, int resident, int sync_each_step)
{
	int s, k;
	Dataset *data = &net->test;
	if (data->input == NULL) { printf("\nERROR: no TEST dataset defined\n"); exit(EXIT_FAILURE); }
	if (resident) dataset_upload(net, data);
	stage_invalidate(net);
	net->is_inference = 1;
	net->length = net->batch_size;
	for (s = 0; s < nsteps; s++) {
		int j = s % data->nb_batch;
		if (!resident) {
			int slot = stage_acquire(net, data->input[j], NULL, s + 1 < nsteps ? data->input[(s + 1) % data->nb_batch] : NULL, NULL);
			use_device_batch(net, net->stage_in[slot]);
		} else use_device_batch(net, data->input_device[j]);
		for (k = 0; k < net->nb_layers; k++) net->net_layers[k]->forward(net->net_layers[k]);
		if (!resident) {
			/* device->host read of the step's result: the class scores of the batch */
			layer *last = net->net_layers[net->nb_layers - 1];
			size_t bytes = (size_t)net->batch_size * last->out_h * last->out_w * cb200_round_channels(last->out_c) * cb200_dtype_size(net->dtype);
			if (net->out_host == NULL) CB_CHECK(cb200_host_alloc(&net->out_host, bytes));
			CB_CHECK(cb200_d2h(net->out_host, last->output, bytes, NULL));
		}
		if (sync_each_step) CB_CHECK(cb200_stream_sync(NULL));
	}
}

static void progress(network *net, int done, int total, double loss, double ips)
{
	int i, size = net->adv_size, filled = (int)((double)done / total * size);
	printf("\r[");
	for (i = 0; i < size; i++) printf(i < filled ? "#" : "-");
	printf("] %d/%d  Loss: %.5g  it/s: %.1f ", done, total, loss, ips);
	fflush(stdout);
}

void train_network(network *net, int nb_iter, int control_interv, float u_begin_learning_rate, float u_end_learning_rate, float u_momentum,
	float u_decay, float u_weight_decay, int show_confmat, int save_every, int save_bin, int shuffle_gpu, int shuffle_every, float c_TC_scale_factor, int silent)
{
	int i, j, k;
	char name[200];

	if (net->inference_only) {
		printf("\n Network was loaded in inference only mode. \n Re-init network with inference only set to false to re-eanble training capability.\n");
		return;
	}
	if (net->train.input == NULL) { printf("\nERROR: no TRAIN dataset defined\n"); exit(EXIT_FAILURE); }
	/* loss scaling is only honoured by the FP16 mode (src/cuda/cuda_main.cu:63-76) */
	net->TC_scale_factor = net->use_cuda_TC == FP16C_FP32A ? c_TC_scale_factor : 1.0f;
	net->momentum = u_momentum; net->decay = u_decay; net->weight_decay = u_weight_decay;
	{
		layer *last = net->net_layers[net->nb_layers - 1];
		net->out_size = last->type == DENSE ? last->out_c + 1 : last->out_c * last->out_h * last->out_w;
		if (last->type == DENSE && net->out_size != net->output_dim + 1) { printf("\nERROR: last layer size does not match the expected output dimensions.\n"); exit(EXIT_FAILURE); }
	}
	prepare_training(net);
	if (!net->dynamic_load) dataset_upload(net, &net->train);
	if (net->iter == 0) remove("error.txt");

	for (i = 0; i < nb_iter; i++) {
		double total_error = 0.0, t_epoch = now_s();
		float lr = u_end_learning_rate + (u_begin_learning_rate - u_end_learning_rate) * expf(-net->decay * net->iter);
		if (silent < 1) printf("\n");
		int shuffled = 0, next_epoch_first = -1, batch_loc = 0, sgd_next = 0;
		net->iter++;
		if (shuffle_every > 0 && (net->iter + 1) % shuffle_every == 0 && net->batch_param != SGD) {
			/* no copy of the previous epoch may still be reading the host batches */
			if (net->copy_stream != NULL) CB_CHECK(cb200_stream_sync(net->copy_stream));
			CB_CHECK(cb200_stream_sync(NULL));
			/* as upstream (src/auxil.c:1768-1785): a resident set is permuted on the device when shuffle_gpu is set,
			 * through the host otherwise; a dynamically loaded set only exists on the host */
			if (shuffle_gpu && !net->dynamic_load) shuffle_dataset_device(net, &net->train);
			else shuffle_dataset(net, &net->train);
			shuffled = 1;
		}
		set_hyper(net, lr, net->momentum, net->weight_decay);
		/* staged copies survive into the next epoch of this call when the host batches cannot have changed in between;
		 * the last batch of such an epoch prefetches the first batch of the next one */
		if (i == 0 || shuffled) stage_invalidate(net);
		if (i + 1 < nb_iter && !(shuffle_every > 0 && (net->iter + 2) % shuffle_every == 0 && net->batch_param != SGD))
			next_epoch_first = 0;
		net->is_inference = 0;
		net->inference_drop_mode = AVG_MODEL;      /* src/auxil.c:1793: the in-training validation pass never draws masks */
		for (j = 0; j < net->train.nb_batch; j++) {
			double t_batch = now_s(), batch_error = 0.0;
			/* perf_eval: the first batch of the first epoch and of every 16th epoch is timed layer by layer, with the weight
			 * gradients on the compute stream so that every kernel lands between its own layer's events */
			const int sample = net->perf_eval && j == 0 && (net->perf_n == 0 || net->iter % 16 == 0);   /* (a sample, not a census) */
			void *side = net->wgrad_stream;
			if (sample) { perf_begin_sample(net); net->wgrad_stream = NULL; }
			/* batch_size == 1 ("SGD" scheme): a random sample per step instead of a sweep (src/auxil.c:1806-1809); the draw for
			 * the next step is made now so that its host->device copy can still be prefetched */
			if (net->batch_param == SGD) {
				if (j == 0) sgd_next = (int)((rand() / ((double)RAND_MAX + 1.0)) * net->train.size);
				batch_loc = sgd_next;
				sgd_next = (int)((rand() / ((double)RAND_MAX + 1.0)) * net->train.size);
				train_one_batch(net, &net->train, batch_loc, !net->dynamic_load, j + 1 < net->train.nb_batch ? sgd_next : -1);
			} else
			train_one_batch(net, &net->train, j, !net->dynamic_load, j + 1 < net->train.nb_batch ? j + 1 : next_epoch_first);
			CB_CHECK(cb200_stream_sync(NULL));
			if (sample) { net->wgrad_stream = side; perf_end_sample(net); }
			for (k = 0; k < net->length; k++) { batch_error += net->loss_host[k]; total_error += net->loss_host[k]; }
			batch_error /= net->length;
			if (isnan(batch_error)) { printf("\nERROR: Network divergence detected (Nan)!\n\n"); exit(EXIT_FAILURE); }
			net->last_batch_loss = (float)batch_error;
			if (silent < 1) progress(net, j + 1, net->train.nb_batch, batch_error, net->batch_size / (now_s() - t_batch));
		}
		net->last_items_per_s = (float)(net->train.size / (now_s() - t_epoch));
		net->last_epoch_loss = total_error / net->train.size;
		if (control_interv > 0 && net->iter % control_interv == 0) {
			if (silent < 1) {
				printf("\n%*s", 14, " ");
				printf("Average Training perf: %0.2f it/s |", net->last_items_per_s);
				printf(" Mean Loss: %.5g |", net->last_epoch_loss);
				printf(" Learning rate: %.5g | Momentum: %.5g | Weight decay: %.5g\n", lr, net->momentum, net->weight_decay);
			}
			if (net->valid.input != NULL) {
				net->is_inference = 1; net->no_error = 0;
				compute_error(net, net->valid, 0, show_confmat, 1, silent);
			}
		}
		if (save_every > 0 && net->iter % save_every == 0) {
			sprintf(name, "net_save/net%d_s%04d.dat", net->id, net->iter);
			printf("Saving network for iteration: %d (mode: %d)\n", net->iter, save_bin);
			save_network(net, name, save_bin);
		}
	}
}

/* ------------------------------------------------------------------ inference */
void compute_error(network *net, Dataset data, int saving, int confusion_matrix, int repeat, int silent)
{
	int j, k;
	double total_error = 0.0, t0 = now_s();
	int c, h, w;
	size_t out_elems;
	float *out_dev = NULL, *out_host = NULL;
	/* Results travel through TWO sets of pinned host buffers: the host-side half of step s (loss sums, confusion matrix,
	 * fwd_res records - upstream formats every value with fprintf) runs while the device already computes step s + 1
	 * (upstream: compute, blocking copies, then the host loop, src/auxil.c:1228-1400). */
	float *loss_h[2] = {NULL, NULL}, *out_h[2] = {NULL, NULL}, *parts_h[2] = {NULL, NULL}, *monitor_h[2] = {NULL, NULL};
	int *am_h[2] = {NULL, NULL}, slot_len[2] = {0, 0}, pending = -1, step = 0;
	void *slot_ev[2] = {NULL, NULL};
	FILE *f_save = NULL;
	char name[200];
	struct stat st;
	layer *last = net->net_layers[net->nb_layers - 1];
	yolo_param *yolo = last->activation_type == YOLO ? (yolo_param *)last->activ_param : NULL;
	double part_err[6] = {0, 0, 0, 0, 0, 0}, sum_IoU = 0.0, sum_obj = 0.0;
	long nb_IoU = 0, nb_good_IoU = 0;
	/* confusion matrix (src/auxil.c:1137-1146, 1384-1426): classification read-out, i.e. one output per class */
	int o = net->output_dim, *am_dev = NULL;
	double *mat = NULL;

	last_layer_dims(net, &c, &h, &w);
	if (confusion_matrix > 0 && !net->no_error && repeat <= 1 && yolo == NULL && o > 0 && c * h * w == o) {
		mat = (double *)calloc((size_t)o * o, sizeof(double));
		CB_CHECK(cb200_malloc((void **)&am_dev, 2 * (size_t)net->batch_size * sizeof(int)));
		for (k = 0; k < 2; k++) CB_CHECK(cb200_host_alloc((void **)&am_h[k], 2 * (size_t)net->batch_size * sizeof(int)));
	}
	for (k = 0; k < 2; k++) {
		CB_CHECK(cb200_event_create(&slot_ev[k]));
		CB_CHECK(cb200_host_alloc((void **)&loss_h[k], (size_t)net->batch_size * sizeof(float)));
		if (yolo != NULL) {
			CB_CHECK(cb200_host_alloc((void **)&parts_h[k], (size_t)net->batch_size * 6 * sizeof(float)));
			CB_CHECK(cb200_host_alloc((void **)&monitor_h[k], (size_t)net->batch_size * h * w * yolo->nb_box * 2 * sizeof(float)));
		}
	}
	out_elems = last->type == DENSE ? (size_t)net->batch_size * (c + 1) : (size_t)net->batch_size * c * h * w;
	if (saving > 0) {
		if (stat("fwd_res", &st) == -1) mkdir("fwd_res", 0700);
		sprintf(name, "fwd_res/net%d_%04d.dat", net->id, net->iter);
		f_save = fopen(name, saving == 1 ? "w+" : "wb+");
		if (f_save == NULL) { printf("ERROR: cannot open %s\n", name); exit(EXIT_FAILURE); }
		CB_CHECK(cb200_malloc((void **)&out_dev, out_elems * sizeof(float)));
		for (k = 0; k < 2; k++) CB_CHECK(cb200_host_alloc((void **)&out_h[k], out_elems * sizeof(float)));
	}
	if (!net->dynamic_load && data.input_device == NULL) {
		/* `data` is a by-value copy (upstream's signature): make the device-resident copies on the network's own object so
		 * that they are created once and found again on the next call */
		Dataset *own = data.input == net->valid.input ? &net->valid : data.input == net->test.input ? &net->test :
			data.input == net->train.input ? &net->train : NULL;
		if (own != NULL) { dataset_upload(net, own); data = *own; }
		else dataset_upload(net, &data);      /* a caller-owned Dataset: the copy lives as long as the caller keeps `data` */
	}
	if (repeat < 1) repeat = 1;
	net->is_inference = 1;
	stage_invalidate(net);
	for (j = 0; j < data.nb_batch; j++) {
		const void *tgt;
		int r, repeat_start = 0;
		net->length = (j == data.nb_batch - 1 && data.size % net->batch_size > 0) ? data.size % net->batch_size : net->batch_size;
		if (net->dynamic_load) {
			/* batch j + 1 travels host -> device on the copy stream while batch j computes (upstream: blocking copy per batch) */
			int slot = stage_acquire(net, data.input[j], data.target[j],
				j + 1 < data.nb_batch ? data.input[j + 1] : NULL, j + 1 < data.nb_batch ? data.target[j + 1] : NULL);
			use_device_batch(net, net->stage_in[slot]);
			tgt = net->stage_tg[slot];
		} else {
			use_device_batch(net, data.input_device[j]);
			tgt = data.target_device[j];
		}
		/* MC-dropout: `repeat` forward passes per batch, restarted from the first layer that has dropout (everything
		 * below it is deterministic), every pass saved and counted in the loss (src/auxil.c:1216-1226) */
		for (r = 0; r <= repeat; r++) {
		/* ---- host half of the step enqueued one turn ago (its results are in slot `pending`), while the device works on
		 * the step just enqueued; the turn r == repeat of the last batch only drains */
		const int drain_only = r == repeat;
		int slot = step & 1;
		if (drain_only && j + 1 < data.nb_batch) break;
		if (!drain_only) {
		for (k = repeat_start; k < net->nb_layers; k++) {
			if (repeat_start == 0 && net->net_layers[k]->dropout_rate > 0.01f) repeat_start = k;
			net->net_layers[k]->forward(net->net_layers[k]);
		}
		if (!net->no_error) {
			output_error(net, tgt);
			CB_CHECK(cb200_d2h(loss_h[slot], net->loss_dev, (size_t)net->batch_size * sizeof(float), NULL));
			if (yolo != NULL) {
				CB_CHECK(cb200_d2h(parts_h[slot], yolo->parts_dev, (size_t)net->batch_size * 6 * sizeof(float), NULL));
				CB_CHECK(cb200_d2h(monitor_h[slot], yolo->monitor_dev, (size_t)net->batch_size * h * w * yolo->nb_box * 2 * sizeof(float), NULL));
			}
		}
		if (mat != NULL) {
			CB_CHECK(cb200_output_argmax(am_dev, am_dev + net->batch_size, last->output, tgt, net->dtype, net->batch_size, net->length, c, h, w, NULL));
			CB_CHECK(cb200_d2h(am_h[slot], am_dev, 2 * (size_t)net->batch_size * sizeof(int), NULL));
		}
		if (saving > 0) {
			if (yolo != NULL && net->y_param->raw_output == 0) CB_CHECK(cb200_yolo_export_boxes(&yolo->desc, out_dev, last->output, NULL));
			else if (last->type == DENSE) CB_CHECK(cb200_export_dense(out_dev, last->output, net->dtype, net->batch_size, c, 0.0f, NULL));
			else CB_CHECK(cb200_export_cbhw(out_dev, last->output, net->dtype, net->batch_size, c, h, w, NULL));
			CB_CHECK(cb200_d2h(out_h[slot], out_dev, out_elems * sizeof(float), NULL));
		}
		CB_CHECK(cb200_event_record(slot_ev[slot], NULL));
		slot_len[slot] = net->length;
		step++;
		}
		if (pending >= 0) {
		const int len = slot_len[pending];
		const float *loss_host = loss_h[pending], *parts_host = parts_h[pending], *monitor_host = monitor_h[pending];
		const int *am_host = am_h[pending];
		out_host = out_h[pending];
		CB_CHECK(cb200_event_sync(slot_ev[pending]));
		if (!net->no_error) for (k = 0; k < len; k++) total_error += loss_host[k];
		if (mat != NULL)
			for (k = 0; k < len; k++) {
				const int truth = am_host[net->batch_size + k], pred = am_host[k];
				if (truth >= 0 && truth < o && pred >= 0 && pred < o) mat[(size_t)truth * o + pred] += 1.0;
			}
		if (!net->no_error && yolo != NULL) {
			/* loss split and association statistics of the batch (src/auxil.c:1429-1486) */
			size_t m, nm = (size_t)net->batch_size * h * w * yolo->nb_box;
			for (k = 0; k < len * 6; k++) part_err[k % 6] += parts_host[k];
			for (m = 0; m < nm; m++)
				if (monitor_host[2 * m] > -0.98f) {
					nb_IoU++;
					sum_obj += monitor_host[2 * m];
					sum_IoU += monitor_host[2 * m + 1];
					if (monitor_host[2 * m + 1] >= net->y_param->IoU_limits[0]) nb_good_IoU++;
				}
		}
		if (saving > 0) {
			/* one line / record per sample, sample-major like upstream's fwd_res files (src/auxil.c:1346-1400); with
			 * repeat > 1 the file holds, batch after batch, `repeat` consecutive blocks of the batch's samples */
			/* a dense output record is out_size = nb_neurons + 1 values: upstream writes the bias node too (src/auxil.c:1262-1276) */
			int b, o, per = last->type == DENSE ? c + 1 : c * h * w;
			for (b = 0; b < len; b++) {
				for (o = 0; o < per; o++) {
					float v = last->type == DENSE ? out_host[(size_t)b * (c + 1) + o]
						: out_host[((size_t)(o / (h * w)) * net->batch_size + b) * (h * w) + o % (h * w)];
					if (saving == 1) fprintf(f_save, "%g ", v); else fwrite(&v, sizeof(float), 1, f_save);
				}
				if (saving == 1) fprintf(f_save, "\n");
			}
		}
		}
		pending = drain_only ? -1 : slot;
		}
	}
	CB_CHECK(cb200_stream_sync(NULL));
	net->last_items_per_s = (float)(data.size / (now_s() - t0));
	/* mean over samples AND repeats on the screen (src/auxil.c:1506-1513); error.txt keeps upstream's total / data.size */
	net->last_epoch_loss = data.size > 0 ? total_error / ((double)data.size * repeat) : 0.0;
	if (!net->no_error && isnan(total_error)) { printf("\nERROR: Network divergence detected (Nan)!\n\n"); exit(EXIT_FAILURE); }
	if (silent < 1) {
		const double norm = (double)data.size * repeat;
		printf("\n%*s", 14, " ");
		printf("Average forward perf : %0.2f it/s ", net->last_items_per_s);
		if (!net->no_error) printf("| Mean Loss: %.5g", net->last_epoch_loss);
		if (!net->no_error && yolo != NULL && data.size > 0)
			printf("\nLoss dist. ||Pos: %.5f |Size: %.5f |Prob: %.5f |Obj: %.5f |Class: %.5f |Param: %.5f ||M IoU = %.4f |M Obj = %0.4f |P Good = %0.4f",
				part_err[0] / norm, part_err[1] / norm, part_err[2] / norm, part_err[3] / norm, part_err[4] / norm,
				part_err[5] / norm, sum_IoU / nb_IoU, sum_obj / nb_IoU, (float)nb_good_IoU / (float)nb_IoU);
		printf("\n");
	}
	if (net->no_error == 0 && silent != 1) {
		/* learning-curve file, same columns as upstream (src/auxil.c:1528-1547) */
		FILE *f_err = fopen("error.txt", "a");
		if (f_err != NULL) {
			fprintf(f_err, "%d %g", net->iter, data.size > 0 ? total_error / data.size : 0.0);
			if (yolo != NULL && data.size > 0)
				fprintf(f_err, " %g %g %g %g %g %g", part_err[0] / data.size, part_err[1] / data.size, part_err[2] / data.size,
					part_err[3] / data.size, part_err[4] / data.size, part_err[5] / data.size);
			fprintf(f_err, "\n");
			fclose(f_err);
		}
	}
	if (f_save != NULL) { fclose(f_save); cb200_free(out_dev); }
	for (k = 0; k < 2; k++) {
		cb200_host_free(loss_h[k]); cb200_host_free(out_h[k]); cb200_host_free(parts_h[k]); cb200_host_free(monitor_h[k]);
		cb200_event_destroy(slot_ev[k]);
	}
	if (mat != NULL) {
		/* same three report levels as upstream (src/auxil.c:1562-1652): rows = true class, columns = predicted class */
		double *recall = (double *)calloc(o, sizeof(double)), *prec = (double *)calloc(o, sizeof(double)), count = 0.0;
		int a, b;
		for (a = 0; a < o; a++) {
			double row = 0.0, col = 0.0;
			for (b = 0; b < o; b++) { row += mat[(size_t)a * o + b]; col += mat[(size_t)b * o + a]; }
			recall[a] = mat[(size_t)a * o + a] / row * 100.0;
			prec[a] = mat[(size_t)a * o + a] / col * 100.0;
			count += mat[(size_t)a * o + a];
		}
		net->last_accuracy = data.size > 0 ? count / data.size : 0.0;
		if (silent != 1) {
			if (confusion_matrix == 1) {
				int width = (o * 10) / 2;
				printf("\n   ");
				for (a = 0; a < width - 3; a++) printf("*");
				printf("  ConfMat  ");
				for (a = 0; a < width - 3; a++) printf("*");
				printf("   Recall\n");
				for (a = 0; a < o; a++) {
					printf("%*s", 5, " ");
					for (b = 0; b < o; b++) printf("%8d |", (int)mat[(size_t)a * o + b]);
					printf("%11.2f%%\n", recall[a]);
				}
				printf("%6s", "Prec. ");
				for (a = 0; a < o; a++) printf("%7.2f%%  ", prec[a]);
				printf("Acc %6.2f%%\n", count / data.size * 100);
			} else if (confusion_matrix == 2) {
				printf("\n   \n Recall:   ");
				for (a = 0; a < o; a++) printf("%7.2f%%  ", recall[a]);
				printf("\n Precision:");
				for (a = 0; a < o; a++) printf("%7.2f%%  ", prec[a]);
				printf("\n Accuracy: %6.2f%%\n", count / data.size * 100);
			} else
				printf("\n Accuracy: %6.2f%%\n", count / data.size * 100);
		}
		free(recall); free(prec); free(mat);
		cb200_free(am_dev); cb200_host_free(am_h[0]); cb200_host_free(am_h[1]);
	}
}

void forward_testset(network *net, int saving, int repeat, int drop_mode, int silent)
{
	if (net->test.input == NULL) { printf("\nERROR: no TEST dataset defined\n"); exit(EXIT_FAILURE); }
	if (repeat > 1 && silent != 1) printf("Forwarding with repeat = %d\n", repeat);
	net->inference_drop_mode = drop_mode;
	net->is_inference = 1;
	compute_error(net, net->test, saving, 0, repeat, silent);
}

/* ------------------------------------------------------------------ checkpoint I/O */
void save_network(network *net, const char *filename, int f_bin)
{
	int i;
	FILE *f;
	struct stat st;
	if (stat("net_save", &st) == -1) mkdir("net_save", 0700);
	f = fopen(filename, f_bin ? "wb+" : "w+");
	if (f == NULL) { printf("ERROR : cannot save %s file\n", filename); exit(EXIT_FAILURE); }
	if (f_bin) fwrite(net->in_dims, sizeof(int), 4, f);
	else fprintf(f, "%dx%dx%dx%d\n", net->in_dims[0], net->in_dims[1], net->in_dims[2], net->in_dims[3]);
	for (i = 0; i < net->nb_layers; i++) {
		switch (net->net_layers[i]->type) {
		case CONV: conv_save(f, net->net_layers[i], f_bin); break;
		case POOL: pool_save(f, net->net_layers[i], f_bin); break;
		case NORM: norm_save(f, net->net_layers[i], f_bin); break;
		case LRN: lrn_save(f, net->net_layers[i], f_bin); break;
		case DENSE: dense_save(f, net->net_layers[i], f_bin); break;
		default: printf("ERROR: layer type cannot be saved\n"); exit(EXIT_FAILURE);
		}
	}
	fclose(f);
}

void load_network(network *net, const char *filename, int iter, int nb_layers, int f_bin)
{
	FILE *f;
	int temp_dim[4], layer_count = 0;
	char layer_type = 'A';
	net->iter = iter;
	net->nb_layers = 0;
	f = fopen(filename, f_bin ? "rb+" : "r+");
	if (f == NULL) { printf(" ERROR: cannot load/find %s file\n", filename); exit(EXIT_FAILURE); }
	if (f_bin) fread(temp_dim, sizeof(int), 4, f);
	else fscanf(f, "%dx%dx%dx%d\n", &temp_dim[0], &temp_dim[1], &temp_dim[2], &temp_dim[3]);
	if (net->in_dims[0] != temp_dim[0] || net->in_dims[1] != temp_dim[1] || net->in_dims[2] != temp_dim[2] || net->in_dims[3] != temp_dim[3]) {
		printf(" WARNING: change in image format !\nLoaded network was trained with : W = %d, H = %d, D = %d, C = %d\n", temp_dim[0], temp_dim[1], temp_dim[2], temp_dim[3]);
		if (net->in_dims[3] != temp_dim[3]) { printf(" ERROR: wrong number of input channel !\n"); exit(EXIT_FAILURE); }
	}
	do {
		if (f_bin) { if (fread(&layer_type, sizeof(char), 1, f) != 1) break; }
		else { if (fscanf(f, "%c", &layer_type) == EOF) break; }
		switch (layer_type) {
		case 'C': conv_load(net, f, f_bin); break;
		case 'P': pool_load(net, f, f_bin); break;
		case 'N': norm_load(net, f, f_bin); break;
		case 'D': dense_load(net, f, f_bin); break;
		case 'L': lrn_load(net, f, f_bin); break;
		case ' ':
		case '\n': layer_count--; break;
		default: printf("ERROR: Layer type not recognized when loading the save model, likely file format error!\n"); exit(EXIT_FAILURE);
		}
		layer_count++;
	} while (nb_layers <= 0 || layer_count < nb_layers);
	fclose(f);
}

void set_frozen_layers(network *net, int *tab, int dim)
{
	int i;
	for (i = 0; i < dim; i++) net->net_layers[tab[i]]->frozen = 1;
}

/* Table of upstream's perf_eval_display (src/auxil.c:802-870): mean time of every layer, forward and backprop (weight
 * update included), from the sampled mini-batches (see train_network). */
void perf_eval_display(network *net)
{
	static const char type_char[5] = { 'C', 'P', 'D', 'N', 'L' };   /* layer_type_enum order */
	double total_fwd = 0.0, total_back = 0.0, total;
	int i;
	if (net->perf_eval == 0) return;
	printf("\nTotal Net. nb weights: %lld\n", net->total_nb_param);
	if (net->perf_n == 0) { printf(" WARNING: no layer was benchmarked yet (the table is sampled while training)\n"); return; }
	for (i = 0; i < net->nb_layers; i++) { total_fwd += net->fwd_perf[i] / net->perf_n; total_back += net->back_perf[i] / net->perf_n; }
	total = total_fwd + total_back;
	printf("\n     Layer  Type       Forward             Backprop             Cumulated\n");
	printf("       N     T      [µs]  /  [%%]         [µs]  /  [%%]         [µs]  /  [%%]\n");
	printf("  -------------------------------------------------------------------\n");
	for (i = 0; i < net->nb_layers; i++) {
		const double f = net->fwd_perf[i] / net->perf_n, b = net->back_perf[i] / net->perf_n;
		printf("   %5d     %c   %8.1f / %4.1f      %8.1f / %4.1f      %8.1f / %4.1f\n", i + 1, type_char[net->net_layers[i]->type],
			f, f / total_fwd * 100.0, b, total_back > 0.0 ? b / total_back * 100.0 : 0.0, f + b, (f + b) / total * 100.0);
	}
	printf("  -------------------------------------------------------------------\n");
	printf("   Total         %8.1f µs          %8.1f µs          %8.1f µs       \n\n", total_fwd, total_back, total);
	printf("  (sampled on %d mini-batch(es); last epoch: %.2f it/s)\n", net->perf_n, net->last_items_per_s);
	fflush(stdout);
}

/* per-layer mean times in microseconds (tests): fwd[nb_layers], back[nb_layers]; returns the number of samples */
int cb_perf_eval_read(network *net, double *fwd, double *back)
{
	int i;
	for (i = 0; i < net->nb_layers && net->perf_n > 0; i++) { fwd[i] = net->fwd_perf[i] / net->perf_n; back[i] = net->back_perf[i] / net->perf_n; }
	return net->perf_n;
}

/* ------------------------------------------------------------------ read-back helpers (tests, parity) */
static void export_act(network *net, layer *l, const void *src, float *dst)
{
	float *tmp = NULL;
	size_t n = l->type == DENSE ? (size_t)net->batch_size * (l->out_c + 1) : (size_t)net->batch_size * l->out_c * l->out_h * l->out_w;
	CB_CHECK(cb200_malloc((void **)&tmp, n * sizeof(float)));
	if (l->type == DENSE) CB_CHECK(cb200_export_dense(tmp, src, net->dtype, net->batch_size, l->out_c, 0.0f, NULL));
	else CB_CHECK(cb200_export_cbhw(tmp, src, net->dtype, net->batch_size, l->out_c, l->out_h, l->out_w, NULL));
	CB_CHECK(cb200_d2h(dst, tmp, n * sizeof(float), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
	cb200_free(tmp);
}

/* A group-norm layer fused with the following max-pool keeps neither its full-resolution output nor its delta
 * (host/layers.c): the read-back helpers rebuild them on demand with the un-fused kernels - the output from the stored
 * input with the CURRENT gamma / beta (so: ask before the optimizer step), the delta from the pool layer's delta + map. */
static void *materialize_fused_norm(network *net, layer *l, int want_delta)
{
	norm_param *np = (norm_param *)l->param;
	layer *pool = np->fused_pool;
	pool_param *pp = (pool_param *)pool->param;
	void *tmp = NULL;
	CB_CHECK(cb200_malloc(&tmp, (size_t)net->batch_size * l->out_h * l->out_w * cb200_round_channels(l->out_c) * cb200_dtype_size(net->dtype)));
	if (!want_delta)
		CB_CHECK(cb200_norm_forward(&np->desc, l->previous->output, tmp, np->gamma, np->beta, np->mean, np->var, np->workspace, NULL));
	else
		CB_CHECK(cb200_pool_backward(&pp->desc, pool->delta_o, pp->pool_map, tmp, NULL, NULL, NULL));
	return tmp;
}

static void export_maybe_fused(network *net, int l, float *dst, int want_delta)
{
	layer *cur = net->net_layers[l];
	if (cur->type == NORM && ((norm_param *)cur->param)->fused_pool != NULL) {
		void *tmp = materialize_fused_norm(net, cur, want_delta);
		export_act(net, cur, tmp, dst);
		cb200_free(tmp);
	} else
		export_act(net, cur, want_delta ? cur->delta_o : cur->output, dst);
}

void cb_layer_export_output(network *net, int l, float *dst) { export_maybe_fused(net, l, dst, 0); }
void cb_layer_export_delta(network *net, int l, float *dst) { export_maybe_fused(net, l, dst, 1); }

/* the 0/1 dropout mask layer l used in its last forward pass, in the layout of cb_layer_export_output (parity tests):
 * masks are never stored, so this re-evaluates the layer's mask function on a tensor of ones */
void cb_layer_export_dropout_mask(network *net, int l, float *dst)
{
	layer *cur = net->net_layers[l];
	size_t count = (size_t)net->batch_size * cur->out_h * cur->out_w * cb200_round_channels(cur->out_c), i;
	size_t es = cb200_dtype_size(net->dtype);
	void *host = malloc(count * es), *tmp = NULL;
	cb200_dropout_desc d = cur->drop;
	if (!(cur->dropout_rate > 0.01f)) { printf("ERROR: layer %d has no dropout\n", l); exit(EXIT_FAILURE); }
	for (i = 0; i < count; i++) {
		if (es == 4) ((float *)host)[i] = 1.0f;
		else ((uint16_t *)host)[i] = net->dtype == CB200_FP16 ? 0x3C00 : 0x3F80;
	}
	CB_CHECK(cb200_malloc(&tmp, count * es));
	CB_CHECK(cb200_h2d(tmp, host, count * es, NULL));
	d.activ.type = CB200_LINEAR; d.length = net->batch_size;
	CB_CHECK(cb200_dropout_forward(&d, tmp, 0, NULL));
	export_act(net, cur, tmp, dst);
	cb200_free(tmp);
	free(host);
}

void cb_layer_export_pool_map(network *net, int l, int *dst)
{
	layer *cur = net->net_layers[l];
	pool_param *p = (pool_param *)cur->param;
	int32_t *tmp = NULL;
	size_t n = (size_t)net->batch_size * cur->out_c * cur->out_h * cur->out_w;
	if (cur->type != POOL || p->pool_map == NULL) { printf("ERROR: layer %d has no pool map\n", l); exit(EXIT_FAILURE); }
	CB_CHECK(cb200_malloc((void **)&tmp, n * sizeof(int32_t)));
	CB_CHECK(cb200_export_pool_map(tmp, p->pool_map, net->batch_size, cur->out_c, cur->out_h, cur->out_w, NULL));
	CB_CHECK(cb200_d2h(dst, tmp, n * sizeof(int32_t), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
	cb200_free(tmp);
}

void cb_layer_shape(network *net, int l, int *out4)
{
	layer *cur = net->net_layers[l];
	out4[0] = cur->out_c; out4[1] = cur->out_h; out4[2] = cur->out_w; out4[3] = cur->type;
}

size_t cb_layer_weight_count(network *net, int l)
{
	layer *cur = net->net_layers[l];
	if (cur->type == CONV) return cb200_conv_master_elems(&((conv_param *)cur->param)->desc);
	if (cur->type == DENSE) { dense_param *p = (dense_param *)cur->param; return (size_t)p->in_size * (p->nb_neurons + 1); }
	if (cur->type == NORM) return 2 * (size_t)((norm_param *)cur->param)->nb_group;
	return 0;
}

extern void dense_get_weights(layer *cur, float *dst, int moment);
extern void dense_set_weights(layer *cur, const float *src);

void cb_layer_get_weights(network *net, int l, float *dst)
{
	layer *cur = net->net_layers[l];
	if (cur->type == CONV) {
		conv_param *p = (conv_param *)cur->param;
		CB_CHECK(cb200_d2h(dst, p->w.master, cb200_conv_master_elems(&p->desc) * sizeof(float), NULL));
	} else if (cur->type == NORM) {
		norm_param *p = (norm_param *)cur->param;
		CB_CHECK(cb200_d2h(dst, p->gamma, p->nb_group * sizeof(float), NULL));
		CB_CHECK(cb200_d2h(dst + p->nb_group, p->beta, p->nb_group * sizeof(float), NULL));
	} else if (cur->type == DENSE) {
		dense_get_weights(cur, dst, 0);
	}
	CB_CHECK(cb200_stream_sync(NULL));
}

void cb_layer_get_moment(network *net, int l, float *dst)
{
	layer *cur = net->net_layers[l];
	if (cur->type == CONV) {
		conv_param *p = (conv_param *)cur->param;
		CB_CHECK(cb200_d2h(dst, p->w.moment, cb200_conv_master_elems(&p->desc) * sizeof(float), NULL));
	} else if (cur->type == NORM) {
		norm_param *p = (norm_param *)cur->param;
		CB_CHECK(cb200_d2h(dst, p->gamma_update, p->nb_group * sizeof(float), NULL));
		CB_CHECK(cb200_d2h(dst + p->nb_group, p->beta_update, p->nb_group * sizeof(float), NULL));
	} else if (cur->type == DENSE) {
		dense_get_weights(cur, dst, 1);
	}
	CB_CHECK(cb200_stream_sync(NULL));
}

void cb_layer_set_weights(network *net, int l, const float *src)
{
	layer *cur = net->net_layers[l];
	if (cur->type == CONV) {
		conv_param *p = (conv_param *)cur->param;
		CB_CHECK(cb200_h2d(p->w.master, src, cb200_conv_master_elems(&p->desc) * sizeof(float), NULL));
		CB_CHECK(cb200_stream_sync(NULL));
		CB_CHECK(cb200_conv_prepare_weights(&p->desc, &p->w, NULL));
	} else if (cur->type == NORM) {
		norm_param *p = (norm_param *)cur->param;
		CB_CHECK(cb200_h2d(p->gamma, src, p->nb_group * sizeof(float), NULL));
		CB_CHECK(cb200_h2d(p->beta, src + p->nb_group, p->nb_group * sizeof(float), NULL));
	} else if (cur->type == DENSE) {
		dense_set_weights(cur, src);
	}
	CB_CHECK(cb200_stream_sync(NULL));
}

void cb_layer_get_norm_stats(network *net, int l, float *mean, float *var, float *d_gamma, float *d_beta)
{
	layer *cur = net->net_layers[l];
	norm_param *p = (norm_param *)cur->param;
	size_t n = (size_t)p->nb_group * net->batch_size * sizeof(float);
	if (cur->type != NORM) { printf("ERROR: layer %d is not a norm layer\n", l); exit(EXIT_FAILURE); }
	if (mean) CB_CHECK(cb200_d2h(mean, p->mean, n, NULL));
	if (var) CB_CHECK(cb200_d2h(var, p->var, n, NULL));
	if (d_gamma && p->d_gamma) CB_CHECK(cb200_d2h(d_gamma, p->d_gamma, n, NULL));
	if (d_beta && p->d_beta) CB_CHECK(cb200_d2h(d_beta, p->d_beta, n, NULL));
	CB_CHECK(cb200_stream_sync(NULL));
}

/* ------------------------------------------------------------------ accessors for the language bindings */
network *cb_get_network(int id) { return (id >= 0 && id < MAX_NETWORKS_NB) ? networks[id] : NULL; }
int cb_net_nb_layers(network *net) { return net->nb_layers; }
int cb_nb_networks(void) { return nb_networks; }
layer *cb_net_layer(network *net, int idx) { return (idx >= 0 && idx < net->nb_layers) ? net->net_layers[idx] : NULL; }
int cb_net_batch_size(network *net) { return net->batch_size; }
float cb_net_last_items_per_s(network *net) { return net->last_items_per_s; }
double cb_net_last_epoch_loss(network *net) { return net->last_epoch_loss; }
double cb_net_last_accuracy(network *net) { return net->last_accuracy; }
void cb_net_set_no_error(network *net, int v) { net->no_error = v; }

Dataset *cb_net_dataset(network *net, const char *name)
{
	if (strcmp(name, "TRAIN") == 0) return &net->train;
	if (strcmp(name, "VALID") == 0) return &net->valid;
	if (strcmp(name, "TEST") == 0) return &net->test;
	if (strcmp(name, "TRAIN_buf") == 0) return &net->train_buf;
	if (strcmp(name, "VALID_buf") == 0) return &net->valid_buf;
	if (strcmp(name, "TEST_buf") == 0) return &net->test_buf;
	printf("ERROR: unknown dataset name %s\n", name);
	exit(EXIT_FAILURE);
}

/* (re)create the named dataset and fill it from row-major FP32 arrays [size][input_dim] / [size][output_dim] */
/* FP32 user arrays -> 16-bit dataset batches with the conversion on the device (upstream converts element by element on
 * the host, src/python_module.c:140-165 + src/cuda/cuda_main.cu:108-113; same round-toward-zero values): one batch of
 * FP32 rows goes up, is packed (cast + bias slot) by cb200_dataset_pack, and comes back into the pinned host batch */
static void dataset_fill_on_device(network *net, Dataset *d, const float *input, const float *target)
{
	size_t es = cb200_dtype_size(net->dtype);
	size_t in_row = net->input_dim + 1, out_row = (size_t)net->output_dim;
	size_t cap = (size_t)net->batch_size * (in_row > out_row ? in_row : out_row);
	float *stage = NULL;
	void *typed = NULL;
	int b;
	CB_CHECK(cb200_malloc((void **)&stage, cap * sizeof(float)));
	CB_CHECK(cb200_malloc(&typed, cap * es));
	for (b = 0; b < d->nb_batch; b++) {
		size_t first = (size_t)b * net->batch_size;
		size_t rows = (size_t)d->size - first < (size_t)net->batch_size ? (size_t)d->size - first : (size_t)net->batch_size;
		CB_CHECK(cb200_h2d(stage, input + first * net->input_dim, rows * net->input_dim * sizeof(float), NULL));
		CB_CHECK(cb200_dataset_pack(typed, net->dtype, stage, rows, net->input_dim, in_row, net->input_bias, NULL));
		CB_CHECK(cb200_d2h(d->input[b], typed, rows * in_row * es, NULL));
		if (target != NULL && out_row > 0) {
			CB_CHECK(cb200_h2d(stage, target + first * out_row, rows * out_row * sizeof(float), NULL));
			CB_CHECK(cb200_dataset_pack(typed, net->dtype, stage, rows, out_row, out_row, 0.0f, NULL));
			CB_CHECK(cb200_d2h(d->target[b], typed, rows * out_row * es, NULL));
		}
	}
	CB_CHECK(cb200_stream_sync(NULL));
	cb200_free(stage); cb200_free(typed);
}

void cb_set_dataset(network *net, const char *name, int size, const float *input, const float *target)
{
	Dataset *d = cb_net_dataset(net, name);
	int i;
	if (d->input != NULL) free_dataset(d);
	*d = create_dataset(net, size);
	if (net->dtype != CB200_FP32 && input != NULL && (size_t)size * net->input_dim >= ((size_t)1 << 18))
		dataset_fill_on_device(net, d, input, target);
	else if (input != NULL || target != NULL)
		for (i = 0; i < size; i++)
			dataset_set_sample(net, d, i, input ? input + (size_t)i * net->input_dim : NULL, target ? target + (size_t)i * net->output_dim : NULL);
	if (!net->dynamic_load) dataset_upload(net, d);
}

/* upstream swap_data_buffers (src/python_module.c:222-283): exchange a dataset with its "_buf" twin */
void cb_swap_data_buffers(network *net, const char *name)
{
	Dataset tmp, *a, *b;
	char buf_name[32];
	snprintf(buf_name, sizeof(buf_name), "%s_buf", name);
	a = cb_net_dataset(net, name);
	b = cb_net_dataset(net, buf_name);
	tmp = *a; *a = *b; *b = tmp;
}
/* ---- YOLO read-backs (bindings, parity tests) */
static yolo_param *yolo_of(network *net)
{
	layer *last = net->net_layers[net->nb_layers - 1];
	if (last->activation_type != YOLO) { printf("ERROR: the last layer is not a YOLO layer\n"); exit(EXIT_FAILURE); }
	return (yolo_param *)last->activ_param;
}
void cb_yolo_set_seed(network *net, unsigned long long seed) { yolo_param *y = yolo_of(net); y->seed = seed; y->step = 0; }
void cb_yolo_box_state(network *net, int *dst)
{
	yolo_param *y = yolo_of(net);
	CB_CHECK(cb200_d2h(dst, y->box_state_dev, (size_t)net->batch_size * y->desc.grid_h * y->desc.grid_w * y->nb_box * sizeof(int), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
}
/* split of the last cb_batch_loss: parts6 = sums over the batch's samples / length, monitor = raw [B][cells][nb_box][2] */
void cb_yolo_loss_parts(network *net, float *parts6, float *monitor)
{
	yolo_param *y = yolo_of(net);
	int k;
	CB_CHECK(cb200_d2h(y->parts_host, y->parts_dev, (size_t)net->batch_size * 6 * sizeof(float), NULL));
	if (monitor != NULL)
		CB_CHECK(cb200_d2h(monitor, y->monitor_dev, (size_t)net->batch_size * y->desc.grid_h * y->desc.grid_w * y->nb_box * 2 * sizeof(float), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
	for (k = 0; k < 6; k++) parts6[k] = 0.0f;
	for (k = 0; k < net->length * 6; k++) parts6[k % 6] += y->parts_host[k] / net->length;
}
void cb_yolo_export_boxes(network *net, float *dst)
{
	yolo_param *y = yolo_of(net);
	layer *last = net->net_layers[net->nb_layers - 1];
	size_t n = (size_t)net->batch_size * last->out_c * last->out_h * last->out_w;
	float *tmp = NULL;
	CB_CHECK(cb200_malloc((void **)&tmp, n * sizeof(float)));
	CB_CHECK(cb200_yolo_export_boxes(&y->desc, tmp, last->output, NULL));
	CB_CHECK(cb200_d2h(dst, tmp, n * sizeof(float), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
	cb200_free(tmp);
}
/* epoch counter (and, when no TRAIN dataset is loaded, the dataset size) seen by the YOLO association pass:
 * it leaves its random start-up phase once iter * train.size > rand_startup (src/activ_functions.c:3006-3015) */
void cb_net_set_iter(network *net, int iter, int train_size)
{
	net->iter = iter;
	if (train_size > 0 && net->train.input == NULL) net->train.size = train_size;
}
void cb_set_wgrad_overlap(network *net, int on)
{
	CB_CHECK(cb200_device_sync());
	if (!on && net->wgrad_stream != NULL) { net->wgrad_stream_off = net->wgrad_stream; net->wgrad_stream = NULL; }
	else if (on && net->wgrad_stream == NULL && net->wgrad_stream_off != NULL) { net->wgrad_stream = net->wgrad_stream_off; net->wgrad_stream_off = NULL; }
}
void cb_net_in_dims(network *net, int *out4) { int i; for (i = 0; i < 4; i++) out4[i] = net->in_dims[i]; }
/* loss scaling is only honoured by the FP16 mode (upstream cuda_set_TC_scale_factor, src/cuda/cuda_main.cu:63-76) */
void cb_set_TC_scale_factor(network *net, float v) { net->TC_scale_factor = net->use_cuda_TC == FP16C_FP32A ? v : 1.0f; }
