function SparseDynamicResid!(T::Vector{<: Real}, residual::AbstractVector{<: Real}, y::Vector{<: Real}, x::Vector{<: Real}, params::Vector{<: Real}, steady_state::Vector{<: Real})
    @assert length(T) >= 11
    @assert length(residual) == 8
    @assert length(y) == 24
    @assert length(x) == 3
    @assert length(params) == 13
@inbounds begin
    residual[1] = T[1] + T[2]*(-T[3] - y[19] - y[20]) - y[17] + y[21] + y[9];
    residual[2] = -T[4]*params[7] - T[5]*y[19] + y[11];
    residual[3] = -T[3] - T[6]*(T[4]*params[4] + params[3]*y[11]) - T[7]*x[2] - params[2]*y[2];
    residual[4] = -T[1] - T[8]*x[1] - params[5]*y[5];
    residual[5] = -T[9]*x[3] - params[6]*y[4] + y[12];
    residual[6] = -params[9] - 100*y[12] + y[14] + 100*y[1] - 100*y[9];
    residual[7] = -params[8] - 400*y[11] + y[15];
    residual[8] = -params[10] - params[8] - 4*params[9] - 400*y[10] + y[16];
end
    return nothing
end
