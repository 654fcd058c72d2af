function SparseDynamicResid!(T::Vector{<: Real}, residual::AbstractVector{<: Real}, y::Vector{<: Real}, x::Vector{<: Real}, params::Vector{<: Real}, steady_state::Vector{<: Real})
    @assert length(T) >= 62
    @assert length(residual) == 16
    @assert length(y) == 48
    @assert length(x) == 2
    @assert length(params) == 7
@inbounds begin
    residual[1] = T[2] + y[30];
    residual[2] = -params[5]*log(y[1]) - x[2] - (1 - params[5])*log(params[4]) + log(y[17]);
    residual[3] = T[10]*T[11]*T[18]*T[9]*params[2]*y[34] - T[8];
    residual[4] = T[20] + y[21];
    residual[5] = -T[20] - T[21]*T[24];
    residual[6] = -T[29] + y[22];
    residual[7] = T[30]*T[31] - T[36]*T[8];
    residual[8] = -T[33] - T[38] + y[19] + y[23];
    residual[9] = T[21] - y[17];
    residual[10] = y[17] + y[24] - y[26] - 1;
    residual[11] = T[39] + y[20];
    residual[12] = -T[33] + y[29];
    residual[13] = -T[41]*y[29] + y[27] + 1;
    residual[14] = -T[45]*y[1] + y[28] + 1;
    residual[15] = y[31] - y[35];
    residual[16] = y[32] - y[34];
end
    return nothing
end
