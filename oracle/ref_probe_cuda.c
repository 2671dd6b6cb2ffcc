/*
 * ref_probe_cuda.c - TEST INFRASTRUCTURE ONLY (oracle side, never linked into the product's kernels).
 *
 * The C_CUDA twin of ref_probe.c: compiled with -D CUDA against the UNMODIFIED reference host sources and
 *   (a) the reference's own src/cuda/ *.cu (nvcc, sm_100)          -> oracle/_ref/cuda/CIANNA.so   (second oracle), or
 *   (b) cianna_b200/shim/cuda_b200_shim.c (the drop-in back-end)   -> oracle/_ref/dropin/CIANNA.so
 * It uses ONLY symbols of the reference's back-end boundary (src/prototypes.h:217-295: cuda_create_table,
 * cuda_put_table, cuda_get_table_to_FP32, cuda_get_table_FP32, ...), so the same driver code runs on either back-end;
 * that is the point of (b).
 *
 * The step driver follows the body of the reference training loop (src/auxil.c:1797-1869).
 */
#include "prototypes.h"

static void *probe_input_dev[MAX_NETWORKS_NB], *probe_target_dev[MAX_NETWORKS_NB], *probe_typed_host[MAX_NETWORKS_NB];
static void *probe_err_dev[MAX_NETWORKS_NB];
static void (*probe_cont_copy[MAX_NETWORKS_NB])(float *elem_in, void *elem_out, int out_offset, size_t nb_elem);
static size_t probe_err_size[MAX_NETWORKS_NB];

int probe_cuda_nb_layers(int net_id) { return networks[net_id]->nb_layers; }
int probe_cuda_layer_type(int net_id, int l) { return networks[net_id]->net_layers[l]->type; }

static size_t typed_size(network *net)
{
	return net->cu_inst.use_cuda_TC == FP32C_FP32A || net->cu_inst.use_cuda_TC == TF32C_FP32A ? 4 : 2;
}

static void ensure_io(network *net)
{
	int id = net->id;
	size_t n_in = (size_t) net->batch_size * (net->input_dim + 1), n_out = (size_t) net->batch_size * net->output_dim;
	if(probe_input_dev[id] != NULL) return;
	cuda_create_table(net, &probe_input_dev[id], n_in);
	cuda_create_table(net, &probe_target_dev[id], n_out);
	probe_typed_host[id] = malloc((n_in > n_out ? n_in : n_out) * typed_size(net));
	{
		Dataset d = cuda_create_dataset(net, 1);   /* only for its cont_copy (FP32 -> compute type, round toward zero) */
		probe_cont_copy[id] = d.cont_copy;
		free(d.input[0]); free(d.target[0]); free(d.input); free(d.target);
	}
}

/* FP32 host rows -> the compute type exactly as the reference fills its datasets (Dataset.cont_copy, round toward zero:
 * src/cuda/cuda_main.cu:108-113,355-371) -> device */
static void put_typed(network *net, void *dev, float *host, size_t n)
{
	probe_cont_copy[net->id](host, probe_typed_host[net->id], 0, n);
	cuda_put_table(net, dev, probe_typed_host[net->id], n);
}

void probe_cuda_forward(int net_id, float *input, int length, int is_inference)
{
	network *net = networks[net_id];
	int k;
	ensure_io(net);
	put_typed(net, probe_input_dev[net_id], input, (size_t) net->batch_size * (net->input_dim + 1));
	net->input = probe_input_dev[net_id];
	net->length = length;
	net->is_inference = is_inference;
	net->inference_drop_mode = AVG_MODEL;
	for(k = 0; k < net->nb_layers; k++)
		net->net_layers[k]->forward(net->net_layers[k]);
}

void probe_cuda_backward(int net_id, float *target, float lr, float momentum, float weight_decay, float TC_scale)
{
	network *net = networks[net_id];
	int k;
	ensure_io(net);
	cuda_set_TC_scale_factor(net, TC_scale);
	put_typed(net, probe_target_dev[net_id], target, (size_t) net->batch_size * net->output_dim);
	net->target = probe_target_dev[net_id];
	net->learning_rate = lr;
	net->momentum = momentum;
	net->weight_decay = weight_decay;
	output_deriv_error(net->net_layers[net->nb_layers-1]);
	for(k = net->nb_layers-1; k >= 0; k--)
		net->net_layers[k]->backprop(net->net_layers[k]);
}

/* per-element loss of the last layer into err[batch_size*out_size] (src/auxil.c:1851-1870) */
void probe_cuda_loss(int net_id, float *target, float *err, int out_size)
{
	network *net = networks[net_id];
	size_t n = (size_t) net->batch_size * out_size, k;
	ensure_io(net);
	put_typed(net, probe_target_dev[net_id], target, (size_t) net->batch_size * net->output_dim);
	net->target = probe_target_dev[net_id];
	net->out_size = out_size;
	if(probe_err_dev[net_id] == NULL || probe_err_size[net_id] != n)
	{
		cuda_create_table_FP32(&probe_err_dev[net_id], n);
		probe_err_size[net_id] = n;
	}
	for(k = 0; k < n; k++) err[k] = 0.0f;
	cuda_put_table_FP32(probe_err_dev[net_id], err, n);
	net->output_error = probe_err_dev[net_id];
	output_error(net->net_layers[net->nb_layers-1]);
	cuda_get_table_FP32(net->output_error, err, n);
}

/* typed activation tensors: what 0 output, 1 delta_o; n elements in the reference layout */
void probe_cuda_read_act(int net_id, int l, int what, float *dst, size_t n)
{
	network *net = networks[net_id];
	layer *cur = net->net_layers[l];
	cuda_get_table_to_FP32(net, what == 0 ? cur->output : cur->delta_o, dst, n, NULL);
}

/* FP32 tables: what 2 master weights (conv: [nb_filters][flat_f_size + TC_padding], dense: [in_size][nb_neurons+1]),
 * 7 mean, 8 var, 9 d_gamma, 10 d_beta (norm, [batch][nb_group] device tables), 5 gamma, 6 beta (host copies upstream
 * keeps current, src/cuda/cuda_norm_layer.cu:455-459) */
void probe_cuda_read_f32(int net_id, int l, int what, float *dst, size_t n)
{
	network *net = networks[net_id];
	layer *cur = net->net_layers[l];
	size_t k;
	switch(cur->type)
	{
		case CONV: {
			conv_param *p = (conv_param*) cur->param;
			if(what == 2) cuda_get_table_FP32(p->FP32_filters, dst, n);
			break; }
		case DENSE: {
			dense_param *p = (dense_param*) cur->param;
			if(what == 2) cuda_get_table_FP32(p->FP32_weights, dst, n);
			break; }
		case NORM: {
			norm_param *p = (norm_param*) cur->param;
			if(what == 5) for(k = 0; k < n; k++) dst[k] = p->gamma[k];
			if(what == 6) for(k = 0; k < n; k++) dst[k] = p->beta[k];
			if(what == 7) cuda_get_table_FP32(p->mean, dst, n);
			if(what == 8) cuda_get_table_FP32(p->var, dst, n);
			if(what == 9) cuda_get_table_FP32(p->d_gamma_gpu, dst, n);
			if(what == 10) cuda_get_table_FP32(p->d_beta_gpu, dst, n);
			break; }
		default: break;
	}
}

/* momentum buffer ("update"): stored in the compute type upstream */
void probe_cuda_read_update(int net_id, int l, float *dst, size_t n)
{
	network *net = networks[net_id];
	layer *cur = net->net_layers[l];
	if(cur->type == CONV) cuda_get_table_to_FP32(net, ((conv_param*) cur->param)->update, dst, n, NULL);
	if(cur->type == DENSE) cuda_get_table_to_FP32(net, ((dense_param*) cur->param)->update, dst, n, NULL);
}

/* overwrite the FP32 master weights (same layouts as probe_cuda_read_f32 what = 2); the 16-bit copy is refreshed by the
 * back-end at the next forward pass (cuda_master_weight_copy, src/cuda/cuda_conv_layer.cu:327) */
void probe_cuda_write_weights(int net_id, int l, float *src, size_t n)
{
	network *net = networks[net_id];
	layer *cur = net->net_layers[l];
	if(cur->type == CONV) cuda_put_table_FP32(((conv_param*) cur->param)->FP32_filters, src, n);
	if(cur->type == DENSE) cuda_put_table_FP32(((dense_param*) cur->param)->FP32_weights, src, n);
}

/* group-norm gamma / beta: upstream keeps host arrays (uploaded at every forward pass, cuda_norm_layer.cu:366-367) */
void probe_cuda_write_norm(int net_id, int l, float *gamma, float *beta, int nb_group)
{
	norm_param *p = (norm_param*) networks[net_id]->net_layers[l]->param;
	int k;
	for(k = 0; k < nb_group; k++) { p->gamma[k] = gamma[k]; p->beta[k] = beta[k]; }
}

void probe_cuda_layer_geom(int net_id, int l, int *out)
{
	layer *cur = networks[net_id]->net_layers[l];
	int i;
	for(i = 0; i < 16; i++) out[i] = 0;
	switch(cur->type)
	{
		case CONV: {
			conv_param *p = (conv_param*) cur->param;
			out[0] = p->nb_filters; out[1] = p->flat_f_size; out[2] = p->TC_padding; out[3] = p->prev_depth;
			for(i = 0; i < 3; i++) { out[4+i] = p->prev_size[i]; out[7+i] = p->nb_area[i]; out[10+i] = p->f_size[i]; }
			break; }
		case POOL: {
			pool_param *p = (pool_param*) cur->param;
			out[0] = p->nb_maps; out[1] = p->pool_type; out[2] = p->global; out[3] = p->prev_depth;
			for(i = 0; i < 3; i++) { out[4+i] = p->prev_size[i]; out[7+i] = p->nb_area[i]; out[10+i] = p->p_size[i]; }
			break; }
		case DENSE: {
			dense_param *p = (dense_param*) cur->param;
			out[0] = p->nb_neurons; out[1] = p->in_size;
			break; }
		case NORM: {
			norm_param *p = (norm_param*) cur->param;
			out[0] = p->n_dim; out[1] = p->group_size; out[2] = p->nb_group; out[3] = p->set_off;
			out[4] = p->dim_offset; out[5] = p->output_dim;
			break; }
		case LRN: {
			lrn_param *p = (lrn_param*) cur->param;
			out[0] = p->n_dim; out[1] = p->range; out[4] = p->dim_offset; out[5] = p->output_dim;
			break; }
		default: break;
	}
}

void probe_cuda_reset(void)
{
	int i;
	for(i = 0; i < MAX_NETWORKS_NB; i++) { probe_input_dev[i] = NULL; probe_target_dev[i] = NULL; probe_err_dev[i] = NULL; }
	nb_networks = 0;
	/* upstream selects the cuBLAS data / compute types once per process (init_cuda, src/cuda/cuda_main.cu:925-1052): a test
	 * process that changes the precision mode between networks has to make it select again */
	is_cuda_init = 0;
}
